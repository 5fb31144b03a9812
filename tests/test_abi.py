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
    tb.write_bundle(prefix, t, with_crc=True)
    back = tb.read_bundle(prefix)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and np.array_equal(back[k], t[k])
    tb.update_checkpoint_state(str(tmp_path / "model"), prefix, [prefix])
    assert tb.latest_checkpoint(str(tmp_path / "model")) == prefix
    assert tb.latest_checkpoint(str(tmp_path)) is None
