"""Reader/writer for TensorFlow "tensor bundle" checkpoints (``<prefix>.index`` +
``<prefix>.data-00000-of-00001``) with no TensorFlow dependency.

The reference saves/restores its model through ``tf.train.Saver``
(reco_utils/recommender/deeprec/models/base_model.py:58, :394-410;
sequential_base_model.py:188-195) and ships a pretrained bundle
(examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz).  This module lets the
B200 build load that bundle and write bundles in the same on-disk format.

Format (restated from the published LevelDB table format + tensor_bundle.proto):
  * ``.index`` is an uncompressed SSTable.  Last 48 bytes = footer: two block handles
    (metaindex, index) as varint (offset, size) pairs, zero padding to 40 bytes, then
    the 8-byte magic 0xdb4775248b80fb57 (little endian).
  * A block of ``size`` bytes = entries, then uint32 restart offsets, then uint32
    num_restarts.  On disk every block is followed by 1 byte (compression type, 0)
    and a 4-byte masked crc32c, not counted in ``size``.
  * Entry = varint shared, varint non_shared, varint value_len, key suffix, value.
  * Index-block values are block handles of data blocks.  Data-block key "" holds a
    BundleHeaderProto; every other key is a tensor name whose value is a
    BundleEntryProto {1: dtype, 2: shape{2: dim{1: size}}, 3: shard_id, 4: offset,
    5: size, 6: crc32c(fixed32)}.
  * ``.data-00000-of-00001`` holds raw little-endian tensor bytes at (offset, size).
"""
import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT_FLOAT, _DT_INT32, _DT_INT64 = 1, 3, 9
_NP_OF_DT = {_DT_FLOAT: np.float32, _DT_INT32: np.int32, _DT_INT64: np.int64, 2: np.float64}
_DT_OF_NP = {np.dtype(np.float32): _DT_FLOAT, np.dtype(np.int32): _DT_INT32,
             np.dtype(np.int64): _DT_INT64, np.dtype(np.float64): 2}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _block_entries(buf, off, size):
    blk = buf[off:off + size]
    (n_restarts,) = struct.unpack_from("<I", blk, size - 4)
    end = size - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        yield key, blk[pos:pos + vlen]
        pos += vlen


def _parse_proto(buf):
    """Minimal protobuf walker -> {field: [values]} (varint ints, bytes, fixed32 ints)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            (v,) = struct.unpack_from("<I", buf, pos)
            pos += 4
        elif wt == 1:
            (v,) = struct.unpack_from("<Q", buf, pos)
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def read_index(prefix, with_header=False):
    """Return {tensor_name: (np_dtype, shape tuple, shard_id, offset, size, masked crc32c)}
    (and the number of data shards from the BundleHeaderProto when ``with_header``)."""
    with open(prefix + ".index", "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise IOError("not a tensor-bundle index: %s.index" % prefix)
    foot = buf[-48:]
    pos = 0
    _, pos = _varint(foot, pos)
    _, pos = _varint(foot, pos)
    idx_off, pos = _varint(foot, pos)
    idx_size, pos = _varint(foot, pos)
    entries = {}
    num_shards = 1
    for _, handle in _block_entries(buf, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _block_entries(buf, boff, bsize):
            if key == b"":
                num_shards = _parse_proto(val).get(1, [1])[0] or 1   # BundleHeaderProto.num_shards
                continue
            pr = _parse_proto(val)
            dtype = pr.get(1, [0])[0]
            dims = []
            if 2 in pr:
                shp = _parse_proto(pr[2][0])
                for d in shp.get(2, []):
                    dims.append(_parse_proto(d).get(1, [0])[0])
            entries[key.decode()] = (
                _NP_OF_DT[dtype], tuple(dims), pr.get(3, [0])[0], pr.get(4, [0])[0], pr.get(5, [0])[0],
                pr.get(6, [0])[0])
    return (entries, num_shards) if with_header else entries


def read_bundle(prefix, names=None, verify_crc=False):
    """Load a checkpoint -> {name: np.ndarray}.  ``prefix`` is e.g. ``.../epoch_3``.  Multi-shard
    bundles (``.data-0000i-of-0000N``) are followed through BundleHeaderProto.num_shards.
    ``verify_crc`` checks every tensor against its stored masked crc32c the way
    BundleReader::GetValue does (IOError on mismatch)."""
    entries, num_shards = read_index(prefix, with_header=True)
    out = {}
    datas = {}
    for name, (dt, shape, shard, off, size, crc) in entries.items():
        if names is not None and name not in names:
            continue
        if shard not in datas:
            path = "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)
            datas[shard] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = np.asarray(datas[shard][off:off + size])
        if verify_crc and _mask(crc32c(raw)) != crc:
            raise IOError("checksum does not match for tensor %s in %s" % (name, prefix))
        out[name] = raw.view(dt).reshape(shape).copy()
    return out


# ---- writer -----------------------------------------------------------------------

def _crc32c_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = None


_NATIVE_CRC = None


def _native_crc():
    """clsr_crc32c of libclsr_b200.so (slice-by-8, ~GB/s) when the library is built; host-only code,
    needs no GPU.  The pure-Python loop below is the same function, ~1 MB/s."""
    global _NATIVE_CRC
    if _NATIVE_CRC is None:
        _NATIVE_CRC = False
        try:
            import ctypes
            from . import engine
            if os.path.exists(engine.LIB_PATH):
                fn = ctypes.CDLL(engine.LIB_PATH).clsr_crc32c
                fn.restype, fn.argtypes = ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
                _NATIVE_CRC = fn
        except (OSError, AttributeError):
            _NATIVE_CRC = False
    return _NATIVE_CRC


def crc32c(data, crc=0):
    """Castagnoli CRC (the one TF stores, masked, per tensor and per block)."""
    fn = _native_crc()
    if fn:
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        return int(fn(a.ctypes.data if a.size else None, a.size, crc))
    return crc32c_py(data, crc)


def crc32c_py(data, crc=0):
    """Pure-Python CRC-32C (reference implementation of ``crc32c``; used when the library is not built)."""
    global _CRC_TAB
    if _CRC_TAB is None:
        _CRC_TAB = np.array(_crc32c_table(), dtype=np.uint32)
    crc ^= 0xFFFFFFFF
    tab = _CRC_TAB
    for b in bytes(data):
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix, tensors, with_crc=True, num_shards=1):
    """Write {name: np.ndarray} as a tensor bundle at ``prefix``.

    Every BundleEntryProto carries the masked crc32c of its tensor bytes: TF's
    BundleReader::GetValue compares it unconditionally, so a bundle without it fails
    ``tf.train.Saver.restore`` with a checksum DataLoss (``with_crc=False`` is for tests only).
    ``num_shards`` > 1 spreads the tensors over ``.data-0000i-of-0000N`` files (sorted names,
    round-robin), the layout a sharded tf.train.Saver produces.
    """
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors)
    num_shards = max(1, int(num_shards))
    items = []
    # BundleHeaderProto: num_shards, endianness little (0), version{producer=1}
    header = b"\x08" + _put_varint(num_shards) + b"\x1a\x02\x08\x01"
    items.append((b"", header))
    files = [open("%s.data-%05d-of-%05d" % (prefix, i, num_shards), "wb") for i in range(num_shards)]
    offs = [0] * num_shards
    try:
        for j, n in enumerate(names):
            sh = j % num_shards
            a = np.ascontiguousarray(tensors[n])
            raw = a.reshape(-1).view(np.uint8) if a.size else np.zeros(0, np.uint8)
            files[sh].write(raw.tobytes() if a.size else b"")
            shape = b"".join(b"\x12" + _put_varint(len(d)) + d
                             for d in (b"\x08" + _put_varint(int(s)) for s in a.shape))
            ent = b"\x08" + _put_varint(_DT_OF_NP[a.dtype])
            ent += b"\x12" + _put_varint(len(shape)) + shape
            if sh:
                ent += b"\x18" + _put_varint(sh)
            if offs[sh]:
                ent += b"\x20" + _put_varint(offs[sh])
            ent += b"\x28" + _put_varint(int(raw.size))
            crc = _mask(crc32c(raw)) if with_crc else 0
            ent += b"\x35" + struct.pack("<I", crc)
            items.append((n.encode(), ent))
            offs[sh] += int(raw.size)
    finally:
        for f in files:
            f.close()
    out = bytearray()

    def emit(block):
        o = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return o, len(block)

    d_off, d_size = emit(_build_block(items))
    m_off, m_size = emit(_build_block([]))
    last_key = items[-1][0] + b"\x00"
    i_off, i_size = emit(_build_block([(last_key, _put_varint(d_off) + _put_varint(d_size))], 1))
    foot = _put_varint(m_off) + _put_varint(m_size) + _put_varint(i_off) + _put_varint(i_size)
    foot = foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", _MAGIC)
    out.extend(foot)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: parse the text ``checkpoint`` state file
    (examples/00_quick_start/sequential.py:352,369)."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                p = line.split(":", 1)[1].strip().strip('"')
                if not os.path.isabs(p):
                    p = os.path.join(model_dir, p)
                return p if os.path.exists(p + ".index") else None
    return None


def update_checkpoint_state(model_dir, prefix, all_prefixes):
    with open(os.path.join(model_dir, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
        for p in all_prefixes:
            f.write('all_model_checkpoint_paths: "%s"\n' % os.path.basename(p))
