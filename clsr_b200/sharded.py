"""Row-sharded embedding tables over NVLink peer memory (BASELINE configs 4-5).

Host side of include/clsr_b200.h's clsr_shard_* entry points: one process per GPU, global row r on
rank r % world at local row r // world.  The reference looks tables up on one device
(sequential_base_model.py:381-437); this replaces that lookup (and its IndexedSlices gradient) when
the item table outgrows one GPU.  torch.distributed is used only to exchange the 64-byte CUDA IPC
handles and for the per-step barriers; rows never travel through a collective."""
import ctypes as C

import numpy as np

from . import engine as E

HANDLE_BYTES = 128  # two cudaIpcMemHandle_t (values, gradient) per rank


def owner_of(ids, world):
    """Rank that stores each global row id."""
    return np.asarray(ids) % world


def local_row_of(ids, world):
    """Row inside the owner's shard."""
    return np.asarray(ids) // world


def local_rows(n_rows, rank, world):
    """Rows of an n_rows table stored on `rank` (ids rank, rank + world, ...)."""
    return max((n_rows - rank + world - 1) // world, 0)


def shard_of(table, rank, world):
    """The rows of a full [n_rows, dim] array that `rank` owns, in local order."""
    return table[rank::world]


def exchange_handles(mine, world, dist=None):
    """All-gather of the per-rank IPC handle blobs, rank-major bytes."""
    if world == 1:
        return bytes(mine)
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(mine))
    assert all(len(g) == HANDLE_BYTES for g in gathered)
    return b"".join(gathered)


class _DevMem:
    """Zero-copy view of engine-owned device memory for torch.as_tensor."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class ShardedTable:
    """One rank's shard of a row-sharded fp32 table (+ gradient shard), with its peers attached."""

    def __init__(self, n_rows, dim, rank=0, world=1, device=0, with_grad=True, dist=None):
        import torch
        self.lib = E.load_library()
        self.n_rows, self.dim, self.rank, self.world, self.device = int(n_rows), int(dim), rank, world, device
        h = C.c_void_p()
        rc = self.lib.clsr_shard_create(device, rank, world, self.n_rows, self.dim, 1 if with_grad else 0, C.byref(h))
        if rc != 0:
            raise E.EngineError("clsr_shard_create: %s" % self.lib.clsr_shard_last_error(None).decode())
        self.h = h
        self.local_rows = int(self.lib.clsr_shard_local_rows(h))
        with torch.cuda.device(device):
            self.values = torch.as_tensor(_DevMem(self.lib.clsr_shard_local_values(h), (self.local_rows, self.dim)),
                                          device="cuda:%d" % device)
            g = self.lib.clsr_shard_local_grad(h)
            self.grad = torch.as_tensor(_DevMem(g, (self.local_rows, self.dim)), device="cuda:%d" % device) if g else None
        blob = (C.c_ubyte * HANDLE_BYTES)()
        self._check(self.lib.clsr_shard_export(h, blob))
        allh = exchange_handles(bytes(blob), world, dist)
        buf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
        self._check(self.lib.clsr_shard_attach(h, buf))

    def _check(self, rc):
        if rc != 0:
            raise E.EngineError(self.lib.clsr_shard_last_error(self.h).decode())

    def load_global(self, full):
        """Fill this rank's shard from a full [n_rows, dim] array-like (numpy or torch, host)."""
        import torch
        part = shard_of(full, self.rank, self.world)
        self.values[:len(part)].copy_(torch.as_tensor(np.ascontiguousarray(part)))

    def zero_grad(self, stream=0):
        self._check(self.lib.clsr_shard_zero_grad(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            self.values = self.grad = None
            self.lib.clsr_shard_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_history(item, cate, item_hist, cate_hist, out=None, stream=0):
    """out[p, :] = concat(item[item_hist[p]], cate[cate_hist[p]]) for int32 CUDA id tensors of any
    shape; returns [*ids.shape, Di + Dc].  One kernel, rows read from the owning GPU over NVLink."""
    import torch
    assert item_hist.dtype == torch.int32 and cate_hist.dtype == torch.int32 and item_hist.is_cuda
    ih, ch = item_hist.contiguous(), cate_hist.contiguous()
    n = ih.numel()
    if out is None:
        out = torch.empty(tuple(ih.shape) + (item.dim + cate.dim,), dtype=torch.float32, device=ih.device)
    item._check(item.lib.clsr_shard_gather_history(item.h, cate.h, ih.data_ptr(), ch.data_ptr(), n, out.data_ptr(),
                                                   C.c_void_p(stream)))
    return out


def scatter_add_history(item, cate, item_hist, cate_hist, d_hist, stream=0):
    """Owner's gradient shard row += d_hist[p, item | cate columns] (16-byte reductions over NVLink)."""
    import torch
    assert d_hist.dtype == torch.float32 and d_hist.is_contiguous()
    ih, ch = item_hist.contiguous(), cate_hist.contiguous()
    item._check(item.lib.clsr_shard_scatter_add_history(item.h, cate.h, ih.data_ptr(), ch.data_ptr(), ih.numel(),
                                                        d_hist.data_ptr(), C.c_void_p(stream)))
