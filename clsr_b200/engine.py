"""ctypes binding of libclsr_b200.so (include/clsr_b200.h) and a thin Python handle.

PyTorch is used only as the owner of device memory (embedding tables and their Adam
slots); every computation happens inside the CUDA library.  There is no CPU fallback:
constructing an Engine without the library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclsr_b200.so")

TABLE_ITEM, TABLE_CATE, TABLE_USER_LONG, TABLE_USER_SHORT = 0, 1, 2, 3
STEP_NO_OPTIMIZER, STEP_NO_BN_UPDATE = 1, 2
EMB = "sequential/embedding/"
TABLE_VARS = {TABLE_ITEM: EMB + "item_embedding", TABLE_CATE: EMB + "cate_embedding",
              TABLE_USER_LONG: EMB + "user_long_embedding", TABLE_USER_SHORT: EMB + "user_short_embedding"}


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("max_rows", C.c_int32), ("seq_len", C.c_int32),
        ("item_dim", C.c_int32), ("cate_dim", C.c_int32), ("user_dim", C.c_int32), ("hidden", C.c_int32),
        ("att0", C.c_int32), ("att1", C.c_int32), ("fc0", C.c_int32), ("fc1", C.c_int32),
        ("n_items", C.c_int64), ("n_cates", C.c_int64), ("n_users", C.c_int64),
        ("train_group", C.c_int32), ("embed_l2", C.c_float), ("layer_l2", C.c_float),
        ("contrastive_kind", C.c_int32), ("triplet_margin", C.c_float), ("contrastive_weight", C.c_float),
        ("discrepancy_weight", C.c_float), ("contrastive_len_threshold", C.c_int32),
        ("contrastive_recent_k", C.c_int32), ("optimizer", C.c_int32), ("learning_rate", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("clip_norm", C.c_int32),
        ("max_grad_norm", C.c_float), ("bn_momentum", C.c_float), ("bn_eps", C.c_float),
        ("math_mode", C.c_int32), ("max_seqs", C.c_int32),
        ("no_interest_evolve", C.c_int32), ("no_predict_long_short", C.c_int32), ("manual_alpha", C.c_int32),
        ("manual_alpha_value", C.c_float), ("sequential_model", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("group", C.c_int32), ("on_device", C.c_int32),
        ("users", C.c_void_p), ("items", C.c_void_p), ("cates", C.c_void_p),
        ("item_history", C.c_void_p), ("cate_history", C.c_void_p), ("mask", C.c_void_p),
        ("time_from_first_action", C.c_void_p), ("time_to_now", C.c_void_p), ("labels", C.c_void_p),
    ]


class Losses(C.Structure):
    _fields_ = [("loss", C.c_float), ("data_loss", C.c_float), ("regular_loss", C.c_float),
                ("contrastive_loss", C.c_float), ("discrepancy_loss", C.c_float),
                ("table_grad_norm", C.c_float * 4)]


EXPORTS = [
    "clsr_create", "clsr_destroy", "clsr_last_error", "clsr_set_stream", "clsr_workspace_bytes",
    "clsr_dense_count", "clsr_dense_name", "clsr_dense_offset", "clsr_dense_numel", "clsr_dense_trainable",
    "clsr_dense_total", "clsr_dense_read", "clsr_dense_write", "clsr_bind_table", "clsr_set_adam_step",
    "clsr_get_adam_step", "clsr_train_step", "clsr_predict", "clsr_synchronize", "clsr_gather_history",
    "clsr_scatter_history_grad", "clsr_sparse_grad_view", "clsr_nccl_unique_id", "clsr_comm_init",
    "clsr_debug_buffer", "clsr_debug_read", "clsr_set_debug_sync", "clsr_kernel_launches", "clsr_set_graphs", "clsr_graph_replays",
    "clsr_set_profiling", "clsr_profile_collect", "clsr_profile_entry", "clsr_debug_gemm", "clsr_debug_dwgemm",
    "clsr_shard_create", "clsr_shard_destroy", "clsr_shard_last_error", "clsr_shard_local_rows",
    "clsr_shard_local_values", "clsr_shard_local_grad", "clsr_shard_export", "clsr_shard_attach",
    "clsr_shard_zero_grad", "clsr_shard_gather_history", "clsr_shard_scatter_add_history",
    "clsr_crc32c", "clsr_clip_report", "clsr_peer_setup_begin", "clsr_peer_setup_finish", "clsr_table_local",
    "clsr_predict_device", "clsr_eval_metrics_compute", "clsr_dataset_create", "clsr_dataset_destroy",
    "clsr_build_batch", "clsr_train_step_staged", "clsr_predict_staged", "clsr_staged_feed_read",
]

_lib = None


def load_library(path=None):
    """dlopen the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python -m clsr_b200.build` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    P, I32, I64, U32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    sig = {
        "clsr_create": (C.c_int, [C.POINTER(Config), C.POINTER(P)]),
        "clsr_destroy": (None, [P]),
        "clsr_last_error": (C.c_char_p, [P]),
        "clsr_set_stream": (C.c_int, [P, P]),
        "clsr_workspace_bytes": (I64, [P]),
        "clsr_dense_count": (I32, [P]),
        "clsr_dense_name": (C.c_char_p, [P, I32]),
        "clsr_dense_offset": (I64, [P, I32]),
        "clsr_dense_numel": (I64, [P, I32]),
        "clsr_dense_trainable": (I32, [P, I32]),
        "clsr_dense_total": (I64, [P]),
        "clsr_dense_read": (C.c_int, [P, I32, P]),
        "clsr_dense_write": (C.c_int, [P, I32, P]),
        "clsr_bind_table": (C.c_int, [P, I32, P, P, P]),
        "clsr_set_adam_step": (C.c_int, [P, I64]),
        "clsr_get_adam_step": (I64, [P]),
        "clsr_train_step": (C.c_int, [P, C.POINTER(Batch), U32, C.POINTER(Losses)]),
        "clsr_predict": (C.c_int, [P, C.POINTER(Batch), P, P]),
        "clsr_synchronize": (C.c_int, [P]),
        "clsr_gather_history": (C.c_int, [P, P, P, I64, P]),
        "clsr_scatter_history_grad": (C.c_int, [P, P, P, I64, P]),
        "clsr_sparse_grad_view": (C.c_int, [P, I32, C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
        "clsr_nccl_unique_id": (C.c_int, [P]),
        "clsr_comm_init": (C.c_int, [P, I32, I32, P]),
        "clsr_debug_buffer": (C.c_int, [P, C.c_char_p, C.POINTER(P), C.POINTER(I64)]),
        "clsr_debug_read": (C.c_int, [P, P, P, I64]),
        "clsr_set_debug_sync": (C.c_int, [P, I32]),
        "clsr_set_graphs": (C.c_int, [P, I32]),
        "clsr_graph_replays": (I64, [P]),
        "clsr_kernel_launches": (I64, [P]),
        "clsr_debug_gemm": (C.c_int, [P, I32, I32, I32, P, I32, P, I32, P, P, I32, I32]),
        "clsr_debug_dwgemm": (C.c_int, [P, I32, I32, I32, P, I32, P, I32, P, I32, P, I32]),
        "clsr_set_profiling": (C.c_int, [P, I32]),
        "clsr_profile_collect": (C.c_int, [P]),
        "clsr_profile_entry": (C.c_int, [P, I32, C.c_char_p, I32, C.POINTER(C.c_double), C.POINTER(I64)]),
        "clsr_shard_create": (C.c_int, [I32, I32, I32, I64, I32, I32, C.POINTER(P)]),
        "clsr_shard_destroy": (None, [P]),
        "clsr_shard_last_error": (C.c_char_p, [P]),
        "clsr_shard_local_rows": (I64, [P]),
        "clsr_shard_local_values": (P, [P]),
        "clsr_shard_local_grad": (P, [P]),
        "clsr_shard_export": (C.c_int, [P, P]),
        "clsr_shard_attach": (C.c_int, [P, P]),
        "clsr_shard_zero_grad": (C.c_int, [P, P]),
        "clsr_shard_gather_history": (C.c_int, [P, P, P, P, I64, P, P]),
        "clsr_shard_scatter_add_history": (C.c_int, [P, P, P, P, I64, P, P]),
        "clsr_crc32c": (U32, [P, C.c_uint64, U32]),
        "clsr_clip_report": (C.c_int, [P, P, C.POINTER(I64)]),
        "clsr_peer_setup_begin": (C.c_int, [P, I32, P]),
        "clsr_peer_setup_finish": (C.c_int, [P, P]),
        "clsr_table_local": (C.c_int, [P, I32, I32, C.POINTER(P), C.POINTER(I64)]),
        "clsr_predict_device": (C.c_int, [P, C.POINTER(Batch), P, P]),
        "clsr_dataset_create": (C.c_int, [P, I64, P, P, P, P, P, P, P, P, P, C.POINTER(P)]),
        "clsr_dataset_destroy": (None, [P]),
        "clsr_build_batch": (C.c_int, [P, P, P, I32, I32, C.c_uint64]),
        "clsr_train_step_staged": (C.c_int, [P, U32, C.POINTER(Losses)]),
        "clsr_predict_staged": (C.c_int, [P, P, P, P, P]),
        "clsr_staged_feed_read": (C.c_int, [P, I32, P, I64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if path == LIB_PATH:
        _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


class DeviceDataset:
    """Handle of a device-resident columnar file cache (clsr_dataset)."""

    def __init__(self, engine, h, n):
        self.engine, self.h, self.n = engine, h, n

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.engine.lib.clsr_dataset_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FEED_KEYS = ("users", "items", "cates", "item_history", "item_cate_history", "mask",
             "time_from_first_action", "time_to_now", "labels")


def normalize_feed(feed, need_labels=True):
    """Cast a feed (placeholder-name -> array) the way TF casts on feed: value-preserving
    float -> int32 for ids / mask (eval batches carry float32 users and mask,
    sequential_iterator.py:670,694), C-contiguous."""
    out = {}
    for k in FEED_KEYS:
        if k == "labels" and (not need_labels or feed.get(k) is None):
            out[k] = None
            continue
        dt = np.float32 if k in ("time_from_first_action", "time_to_now", "labels") else np.int32
        a = np.asarray(feed[k])
        if a.dtype != dt:
            a = a.astype(dt)
        out[k] = np.ascontiguousarray(a.reshape(a.shape[0], -1) if a.ndim > 1 else a)
    return out


def detect_group(feed, group):
    """Rows s*group..s*group+group-1 share user/history/mask/time features?  (True for every
    training batch the reference iterator builds, sequential_iterator.py:588-610.)"""
    if group <= 1:
        return 1
    B = feed["users"].shape[0]
    if B % group:
        return 1
    for k in ("users", "item_history", "item_cate_history", "mask", "time_from_first_action", "time_to_now"):
        a = feed[k].reshape(B // group, group, -1)
        if not (a[:, 1:] == a[:, :1]).all():
            return 1
    return group


class Engine:
    def __init__(self, n_items, n_cates, n_users, max_rows, seq_len=50, item_dim=32, cate_dim=8,
                 user_dim=40, hidden=40, att_sizes=(80, 40), layer_sizes=(100, 64), train_group=5,
                 embed_l2=1e-6, layer_l2=1e-6, contrastive_loss="triplet", triplet_margin=1.0,
                 contrastive_weight=0.1, discrepancy_weight=0.01, contrastive_len_threshold=5,
                 contrastive_recent_k=3, optimizer="adam", learning_rate=1e-3, clip_norm=True,
                 max_grad_norm=2.0, device=0, math_mode=0, training=True, alloc_tables=True, max_seqs=0,
                 interest_evolve=True, predict_long_short=True, manual_alpha=False, manual_alpha_value=0.5,
                 sequential_model="time4lstm"):
        import torch
        if not torch.cuda.is_available():
            raise EngineError("clsr_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.lib = load_library()
        if optimizer not in ("adam", "lazyadam"):
            raise EngineError("optimizer %r is not implemented on the B200 path (adam, lazyadam)" % optimizer)
        if sequential_model not in ("time4lstm", "lstm"):
            raise EngineError("sequential_model %r is not implemented on the B200 path (time4lstm, lstm)" % sequential_model)
        if contrastive_loss not in ("triplet", "bpr"):
            raise EngineError("contrastive_loss %r is not defined (triplet, bpr)" % contrastive_loss)
        self.cfg = Config(
            device=device, max_rows=max_rows, seq_len=seq_len, item_dim=item_dim, cate_dim=cate_dim,
            user_dim=user_dim, hidden=hidden, att0=att_sizes[0], att1=att_sizes[1], fc0=layer_sizes[0],
            fc1=layer_sizes[1], n_items=n_items, n_cates=n_cates, n_users=n_users, train_group=train_group,
            embed_l2=embed_l2, layer_l2=layer_l2, contrastive_kind=1 if contrastive_loss == "bpr" else 0, triplet_margin=triplet_margin,
            contrastive_weight=contrastive_weight, discrepancy_weight=discrepancy_weight,
            contrastive_len_threshold=contrastive_len_threshold, contrastive_recent_k=contrastive_recent_k,
            optimizer=0 if optimizer == "adam" else 1, learning_rate=learning_rate, beta1=0.9, beta2=0.999,
            adam_eps=1e-8, clip_norm=1 if clip_norm else 0, max_grad_norm=max_grad_norm, bn_momentum=0.95,
            bn_eps=1e-4, math_mode=math_mode, max_seqs=int(max_seqs or 0),
            no_interest_evolve=0 if interest_evolve else 1, no_predict_long_short=0 if predict_long_short else 1,
            manual_alpha=1 if manual_alpha else 0, manual_alpha_value=float(manual_alpha_value),
            sequential_model=1 if sequential_model == "lstm" else 0)
        self.device = torch.device("cuda", device)
        self.h = C.c_void_p()
        rc = self.lib.clsr_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            raise EngineError("clsr_create failed (%d): %s" % (rc, self.lib.clsr_last_error(None).decode()))
        n = self.lib.clsr_dense_count(self.h)
        self.layout = [(self.lib.clsr_dense_name(self.h, i).decode(), self.lib.clsr_dense_offset(self.h, i),
                        self.lib.clsr_dense_numel(self.h, i), bool(self.lib.clsr_dense_trainable(self.h, i)))
                       for i in range(n)]
        self.dense_total = self.lib.clsr_dense_total(self.h)
        self.seq_len, self.max_rows = seq_len, max_rows
        self.tables, self.table_m, self.table_v = {}, {}, {}
        shapes = {TABLE_ITEM: (n_items, item_dim), TABLE_CATE: (n_cates, cate_dim),
                  TABLE_USER_LONG: (n_users, user_dim), TABLE_USER_SHORT: (n_users, user_dim)}
        self._n_rows = {t: s[0] for t, s in shapes.items()}
        self.sharded, self.world, self.rank = False, 1, 0
        for t, shp in shapes.items():
            if not alloc_tables:   # tables that only fit row-sharded: comm_init(shard=True) creates the local shards
                continue
            self.tables[t] = torch.zeros(shp, dtype=torch.float32, device=self.device)
            if training:
                self.table_m[t] = torch.zeros(shp, dtype=torch.float32, device=self.device)
                self.table_v[t] = torch.zeros(shp, dtype=torch.float32, device=self.device)
            self._bind(t)
        self.last_table_grad_norms = (0.0, 0.0, 0.0, 0.0)
        self.user_table = None  # sequential/embedding/user_embedding: gathered but unused by CLSR
        # Share torch's current stream so engine kernels are ordered with the torch ops that fill or read
        # the tables and device-resident feeds (the engine's own stream is non-blocking).
        self._check(self.lib.clsr_set_stream(self.h, torch.cuda.current_stream(self.device).cuda_stream))

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise EngineError("clsr_b200 error %d: %s" % (rc, self.lib.clsr_last_error(self.h).decode()))

    def _bind(self, t):
        m = self.table_m.get(t)
        v = self.table_v.get(t)
        self._check(self.lib.clsr_bind_table(self.h, t, self.tables[t].data_ptr(),
                                             m.data_ptr() if m is not None else None,
                                             v.data_ptr() if v is not None else None))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.clsr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- variables --
    def _dense_buf(self, which):
        buf = np.empty(self.dense_total, np.float32)
        self._check(self.lib.clsr_dense_read(self.h, which, buf.ctypes.data))
        return buf

    def _unpack(self, buf, shapes=None):
        out = {}
        for name, off, n, _ in self.layout:
            out[name] = buf[off:off + n].copy()
        return out

    def get_dense(self, which=0):
        """{name: flat float32 array}; which: 0 values, 1 Adam m, 2 Adam v, 3 last gradients."""
        return self._unpack(self._dense_buf(which))

    def set_dense(self, values, which=0, strict=True):
        buf = self._dense_buf(which)
        for name, off, n, _ in self.layout:
            if name in values:
                a = np.asarray(values[name], np.float32).reshape(-1)
                if a.size != n:
                    raise EngineError("variable %s: expected %d values, got %d" % (name, n, a.size))
                buf[off:off + n] = a
            elif strict:
                raise EngineError("missing variable %s" % name)
        self._check(self.lib.clsr_dense_write(self.h, which, buf.ctypes.data))

    def set_params(self, params, strict=True):
        """Load {TF variable name: array} (dense variables + the four trained tables)."""
        self.set_dense(params, 0, strict)
        for t, name in TABLE_VARS.items():
            if name in params:
                src = self.torch.from_numpy(np.ascontiguousarray(params[name], np.float32))
                if self.sharded:   # full table given: keep this rank's rows
                    if src.shape[0] != self._n_rows[t]:
                        raise EngineError("table %s: %d rows != %d" % (name, src.shape[0], self._n_rows[t]))
                    src = src[self.rank::self.world]
                    self.tables[t][:src.shape[0]].copy_(src)
                    continue
                if tuple(src.shape) != tuple(self.tables[t].shape):
                    raise EngineError("table %s: shape %s != %s" % (name, tuple(src.shape), tuple(self.tables[t].shape)))
                self.tables[t].copy_(src)
            elif strict:
                raise EngineError("missing table %s" % name)
        if EMB + "user_embedding" in params:
            self.user_table = np.asarray(params[EMB + "user_embedding"], np.float32)

    def get_params(self):
        out = self.get_dense(0)
        self.synchronize()
        for t, name in TABLE_VARS.items():
            out[name] = self.full_table(t)
        if self.user_table is not None:
            out[EMB + "user_embedding"] = self.user_table
        return out

    def shapes(self):
        """Shapes of the dense variables, by name (from clsr_b200.params)."""
        from . import params as P
        c = self.cfg
        return {n: s for n, s, _, _ in P.dense_spec(c.item_dim + c.cate_dim, c.user_dim, c.hidden,
                                                    [c.att0, c.att1], [c.fc0, c.fc1])}

    # -- steps --
    def _batch(self, feed, group, on_device=False, need_labels=True):
        b = Batch()
        if on_device:
            ptr = lambda k: feed[k].data_ptr() if feed.get(k) is not None else None
            b.rows = int(feed["users"].shape[0])
        else:
            ptr = lambda k: feed[k].ctypes.data if feed.get(k) is not None else None
            b.rows = int(feed["users"].shape[0])
        b.group, b.on_device = group, 1 if on_device else 0
        b.users, b.items, b.cates = ptr("users"), ptr("items"), ptr("cates")
        b.item_history, b.cate_history, b.mask = ptr("item_history"), ptr("item_cate_history"), ptr("mask")
        b.time_from_first_action, b.time_to_now = ptr("time_from_first_action"), ptr("time_to_now")
        b.labels = ptr("labels") if need_labels else None
        return b

    def to_device(self, feed):
        """Upload a normalized feed once (device-resident benchmarking)."""
        t = self.torch
        return {k: (t.from_numpy(v).to(self.device) if v is not None else None) for k, v in feed.items()}

    def train_step(self, feed, group=1, flags=0, on_device=False, normalized=False, wait=True):
        if not on_device and not normalized:
            feed = normalize_feed(feed)
        b = self._batch(feed, group, on_device)
        out = Losses()
        self._keep = feed
        self._check(self.lib.clsr_train_step(self.h, C.byref(b), flags, C.byref(out) if wait else None))
        if not wait:
            return None
        self.last_table_grad_norms = tuple(out.table_grad_norm)
        return {"loss": out.loss, "data_loss": out.data_loss, "regular_loss": out.regular_loss,
                "contrastive_loss": out.contrastive_loss, "discrepancy_loss": out.discrepancy_loss}

    def predict(self, feed, group=1, on_device=False, normalized=False, with_alpha=True):
        if not on_device and not normalized:
            feed = normalize_feed(feed, need_labels=False)
        b = self._batch(feed, group, on_device, need_labels=False)
        pred = np.empty(b.rows, np.float32)
        alpha = np.empty(b.rows, np.float32) if with_alpha else None
        self._check(self.lib.clsr_predict(self.h, C.byref(b), pred.ctypes.data,
                                          alpha.ctypes.data if with_alpha else None))
        return pred, alpha

    def predict_device(self, feed, group=1, on_device=False, normalized=False, with_alpha=False):
        """predict() with the results left on the device: (pred, alpha) float32 CUDA tensors (alpha None unless
        asked for); asynchronous on the engine's stream (= torch's current stream)."""
        if not on_device and not normalized:
            feed = normalize_feed(feed, need_labels=False)
        b = self._batch(feed, group, on_device, need_labels=False)
        t = self.torch
        pred = t.empty(b.rows, dtype=t.float32, device=self.device)
        alpha = t.empty(b.rows, dtype=t.float32, device=self.device) if with_alpha else None
        self._keep = feed
        self._check(self.lib.clsr_predict_device(self.h, C.byref(b), pred.data_ptr(),
                                                 alpha.data_ptr() if with_alpha else None))
        return pred, alpha

    # -- batches built on the device --
    def create_dataset(self, label, user, item, cate, length, item_hist, cate_hist, tfa, ttn):
        """Upload one parsed file (columnar host arrays, histories [n, T] left-aligned) as a device-resident cache."""
        a = lambda x, dt: np.ascontiguousarray(np.asarray(x), dt)
        cols = [a(label, np.float32), a(user, np.int32), a(item, np.int32), a(cate, np.int32), a(length, np.int32),
                a(item_hist, np.int32), a(cate_hist, np.int32), a(tfa, np.float32), a(ttn, np.float32)]
        n = cols[0].shape[0]
        assert cols[5].shape == (n, self.seq_len), cols[5].shape
        h = C.c_void_p()
        self._check(self.lib.clsr_dataset_create(self.h, n, *[c.ctypes.data for c in cols], C.byref(h)))
        return DeviceDataset(self, h, n)

    def build_batch(self, ds, lines, num_ngs, seed=0):
        """Build the feed of the given line indices on the device (num_ngs > 0: training batch with in-batch negatives)."""
        lines = np.ascontiguousarray(lines, np.int32)
        self._check(self.lib.clsr_build_batch(self.h, ds.h, lines.ctypes.data, len(lines), int(num_ngs),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF))
        self._staged = (len(lines), num_ngs + 1)

    def train_step_staged(self, flags=0, wait=True):
        out = Losses()
        self._check(self.lib.clsr_train_step_staged(self.h, flags, C.byref(out) if wait else None))
        if not wait:
            return None
        self.last_table_grad_norms = tuple(out.table_grad_norm)
        return {"loss": out.loss, "data_loss": out.data_loss, "regular_loss": out.regular_loss,
                "contrastive_loss": out.contrastive_loss, "discrepancy_loss": out.discrepancy_loss}

    def predict_staged(self, with_alpha=False, with_users=True):
        """Inference on the staged batch: (pred, alpha, users, labels) CUDA tensors (asynchronous)."""
        t = self.torch
        S, G = self._staged
        B = S * G
        pred = t.empty(B, dtype=t.float32, device=self.device)
        alpha = t.empty(B, dtype=t.float32, device=self.device) if with_alpha else None
        users = t.empty(S, dtype=t.int32, device=self.device) if (with_users and G == 1) else None
        labels = t.empty(B, dtype=t.float32, device=self.device)
        self._check(self.lib.clsr_predict_staged(self.h, pred.data_ptr(), alpha.data_ptr() if with_alpha else None,
                                                 users.data_ptr() if users is not None else None, labels.data_ptr()))
        return pred, alpha, users, labels

    def staged_feed(self):
        """The staged batch as host arrays keyed like a feed (tests)."""
        S, G = self._staged
        T, B = self.seq_len, S * G
        spec = [("item_history", np.int32, (S, T)), ("item_cate_history", np.int32, (S, T)), ("mask", np.int32, (S, T)),
                ("time_from_first_action", np.float32, (S, T)), ("time_to_now", np.float32, (S, T)),
                ("users", np.int32, (S,)), ("items", np.int32, (B,)), ("cates", np.int32, (B,)), ("labels", np.float32, (B,))]
        out = {}
        for i, (k, dt, shp) in enumerate(spec):
            a = np.empty(shp, dt)
            self._check(self.lib.clsr_staged_feed_read(self.h, i, a.ctypes.data, a.nbytes))
            out[k] = a
        return out

    def comm_init(self, rank, world, dist=None, shard=True):
        """Join a data-parallel group of ``world`` engines (one process per GPU).  The NCCL unique id is
        created on rank 0 and broadcast through ``torch.distributed`` (any backend); the CUDA IPC handles of the
        peer-memory communication buffers (and, with ``shard``, of the row-sharded tables) are all-gathered
        the same way.  ``shard=True`` (world a power of two): the four tables are row-sharded over the
        ranks -- rank r keeps global rows r, r + world, ... -- and ``self.tables`` become views of the local shards;
        whatever the full tables held before the call is kept (each rank keeps its rows)."""
        if world <= 1:
            return
        import torch.distributed as td
        dist = dist or td
        buf = C.create_string_buffer(128)
        if rank == 0:
            rc = self.lib.clsr_nccl_unique_id(buf)
            if rc != 0:
                raise EngineError("ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        self._check(self.lib.clsr_comm_init(self.h, rank, world, C.create_string_buffer(box[0], 128)))
        self.world, self.rank = world, rank
        if world & (world - 1):
            shard = False   # round-robin ownership is mask / shift arithmetic: replicated tables otherwise
            return
        blob = C.create_string_buffer(1024)
        self._check(self.lib.clsr_peer_setup_begin(self.h, 1 if shard else 0, blob))
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(blob.raw))
        self._check(self.lib.clsr_peer_setup_finish(self.h, C.create_string_buffer(b"".join(gathered), 1024 * world)))
        if shard:
            self._adopt_shards()

    def _adopt_shards(self):
        """Replace the full-size torch tables by views of the engine-owned local shards (keeping this rank's rows)."""
        t = self.torch
        from .sharded import _DevMem
        self.sharded = True
        dims = {TABLE_ITEM: self.cfg.item_dim, TABLE_CATE: self.cfg.cate_dim, TABLE_USER_LONG: self.cfg.user_dim,
                TABLE_USER_SHORT: self.cfg.user_dim}
        for tb in dims:
            dim = dims[tb]
            views = []
            for which in range(3):
                p, n = C.c_void_p(), C.c_int64()
                self._check(self.lib.clsr_table_local(self.h, tb, which, C.byref(p), C.byref(n)))
                with t.cuda.device(self.device):
                    views.append(t.as_tensor(_DevMem(p.value, (n.value, dim)), device=self.device))
            old = (self.tables.get(tb), self.table_m.get(tb), self.table_v.get(tb))
            for v, o in zip(views, old):
                if o is not None:
                    part = o[self.rank::self.world]
                    v[:part.shape[0]].copy_(part)
            self.tables[tb], self.table_m[tb], self.table_v[tb] = views
        t.cuda.synchronize(self.device)

    def full_table(self, tb, dist=None):
        """The whole [n_rows, dim] table as a numpy array on every rank (collective in sharded mode; tests / checkpoints)."""
        if not getattr(self, "sharded", False):
            return self.tables[tb].cpu().numpy()
        import torch.distributed as td
        dist = dist or td
        self.synchronize()
        parts = [None] * self.world
        dist.all_gather_object(parts, self.tables[tb].cpu().numpy())
        n = self._n_rows[tb]
        out = np.empty((n, parts[0].shape[1]), np.float32)
        for r, p in enumerate(parts):
            rows = len(range(r, n, self.world))
            out[r::self.world] = p[:rows]
        return out

    def synchronize(self):
        self._check(self.lib.clsr_synchronize(self.h))

    def clip_report(self):
        """(table-gradient norms of the last step {table id: norm}, number of shared-history steps so far whose
        table clip was active -- those deviate from TF's clip on un-deduplicated slices; group=1 is exact)."""
        norms = (C.c_float * 4)()
        n = C.c_int64()
        self._check(self.lib.clsr_clip_report(self.h, norms, C.byref(n)))
        return {t: float(norms[t]) for t in range(4)}, int(n.value)

    @property
    def adam_step(self):
        return self.lib.clsr_get_adam_step(self.h)

    @adam_step.setter
    def adam_step(self, v):
        self._check(self.lib.clsr_set_adam_step(self.h, int(v)))

    def kernel_launches(self):
        return int(self.lib.clsr_kernel_launches(self.h))

    def set_graphs(self, on=True):
        """Replay single-GPU training steps as CUDA graphs (default on)."""
        self._check(self.lib.clsr_set_graphs(self.h, 1 if on else 0))

    def graph_replays(self):
        return int(self.lib.clsr_graph_replays(self.h))

    def set_debug_sync(self, on=True):
        self._check(self.lib.clsr_set_debug_sync(self.h, 1 if on else 0))

    def set_stream(self, cuda_stream):
        """Run all engine work on the given CUDA stream (e.g. torch.cuda.Stream().cuda_stream)."""
        self._check(self.lib.clsr_set_stream(self.h, cuda_stream))

    def set_profiling(self, on=True):
        self._check(self.lib.clsr_set_profiling(self.h, 1 if on else 0))

    def profile(self):
        """{kernel name: (total device ms, calls)} since profiling was switched on."""
        n = self.lib.clsr_profile_collect(self.h)
        if n < 0:
            self._check(n)
        out = {}
        buf = C.create_string_buffer(128)
        ms, calls = C.c_double(), C.c_int64()
        for i in range(n):
            self._check(self.lib.clsr_profile_entry(self.h, i, buf, 128, C.byref(ms), C.byref(calls)))
            out[buf.value.decode()] = (ms.value, calls.value)
        return out

    # -- introspection --
    def debug(self, name, shape):
        """Copy a named intermediate buffer of the last step to the host."""
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.clsr_debug_buffer(self.h, name.encode(), C.byref(p), C.byref(n)))
        cnt = int(np.prod(shape))
        if cnt > n.value:
            raise EngineError("buffer %s holds %d floats, asked for %d" % (name, n.value, cnt))
        out = np.empty(cnt, np.float32)
        self._check(self.lib.clsr_debug_read(self.h, p.value, out.ctypes.data, cnt * 4))
        return out.reshape(shape)

    def sparse_grad(self, table):
        """(unique ids, gradient rows) of the last step for one table."""
        ids, rows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.clsr_sparse_grad_view(self.h, table, C.byref(ids), C.byref(rows), C.byref(cnt)))
        n = np.zeros(1, np.int32)
        self._check(self.lib.clsr_debug_read(self.h, cnt.value, n.ctypes.data, 4))
        n = int(n[0])
        dim = int(self.tables[table].shape[1])
        i = np.empty(n, np.int32)
        r = np.empty((n, dim), np.float32)
        if n:
            self._check(self.lib.clsr_debug_read(self.h, ids.value, i.ctypes.data, n * 4))
            self._check(self.lib.clsr_debug_read(self.h, rows.value, r.ctypes.data, n * dim * 4))
        return i, r
