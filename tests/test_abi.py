"""The C-ABI library loads and exports every symbol include/clsr_b200.h declares (no compute)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from clsr_b200 import build, engine
    build.build()
    lib = engine.load_library()
    hdr = open(os.path.join(ROOT, "include", "clsr_b200.h")).read()
    declared = set(re.findall(r"\b(clsr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.clsr_last_error(None) is not None


def test_struct_layouts_match_header():
    """ctypes mirrors of clsr_config / clsr_batch keep the header's field order."""
    from clsr_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "clsr_b200.h")).read()
    start = hdr.index("typedef struct clsr_config {") + len("typedef struct clsr_config {")
    body = hdr[start:hdr.index("} clsr_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(?:int32_t|int64_t|float)\s+(.*)", decl)
        if m:
            fields += [f.strip() for f in m.group(1).split(",")]
    assert fields == [f[0] for f in engine.Config._fields_]


def test_engine_refuses_without_gpu():
    import pytest
    import torch
    from clsr_b200.engine import Engine, EngineError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(EngineError):
        Engine(10, 10, 10, max_rows=10)


def test_tensor_bundle_roundtrip(tmp_path):
    from clsr_b200 import tf_bundle as tb
    rng = np.random.default_rng(0)
    t = {"a/b/kernel": rng.standard_normal((7, 5)).astype(np.float32), "a/b/bias": np.zeros(5, np.float32),
         "a/c": rng.standard_normal((3, 2, 2)).astype(np.float32), "z": np.arange(4, dtype=np.int32)}
    prefix = str(tmp_path / "model" / "epoch_2")
    tb.write_bundle(prefix, t)                       # default: per-tensor crc32c stored (Saver.save path)
    back = tb.read_bundle(prefix, verify_crc=True)   # ... and checked the way BundleReader::GetValue does
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and np.array_equal(back[k], t[k])
    assert all(e[5] != 0 for e in tb.read_index(prefix).values())
    # a flipped data byte is detected
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[5] ^= 0x40
    open(data, "wb").write(bytes(raw))
    import pytest
    with pytest.raises(IOError):
        tb.read_bundle(prefix, verify_crc=True)
    # multi-shard bundles (BundleHeaderProto.num_shards, BundleEntryProto.shard_id)
    p3 = str(tmp_path / "model" / "epoch_3")
    tb.write_bundle(p3, t, num_shards=3)
    assert sorted(f for f in os.listdir(str(tmp_path / "model")) if f.startswith("epoch_3.data")) == \
        ["epoch_3.data-0000%d-of-00003" % i for i in range(3)]
    back3 = tb.read_bundle(p3, verify_crc=True)
    assert all(np.array_equal(back3[k], t[k]) for k in t)
    tb.update_checkpoint_state(str(tmp_path / "model"), prefix, [prefix])
    assert tb.latest_checkpoint(str(tmp_path / "model")) == prefix
    assert tb.latest_checkpoint(str(tmp_path)) is None


def test_crc32c_matches_values_tensorflow_stored():
    """Known answers: the masked crc32c TF wrote into the shipped checkpoint's index for every dense
    variable (tests/golden/ckpt_crc.json, made by tests/golden/make_ckpt_crc.py) against the checksum
    of the same bytes here -- native (clsr_crc32c) and pure-Python implementations."""
    import json
    from clsr_b200 import build, tf_bundle as tb
    build.build()
    assert tb._native_crc(), "clsr_crc32c not loadable"
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ckpt_crc.json")))
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_slice.npz"))
    checked = 0
    for name, g in gold.items():
        if name not in z.files or z[name].nbytes != g["bytes"]:
            continue   # tables are stored sliced
        raw = np.ascontiguousarray(z[name]).reshape(-1).view(np.uint8)
        assert tb._mask(tb.crc32c(raw)) == g["crc32c_masked"], name
        if raw.size <= 4096:
            assert tb.crc32c_py(raw) == tb.crc32c(raw), name
        checked += 1
    assert checked >= 60
    assert tb.crc32c(b"123456789") == 0xE3069283 == tb.crc32c_py(b"123456789")   # CRC-32C check value
    assert tb.crc32c(b"6789", tb.crc32c(b"12345")) == 0xE3069283                  # continuation
