#!/usr/bin/env python
"""Row-sharded history gather + sparse-gradient scatter-add over NVLink peer memory
(BASELINE config 4: 50M-item table, seq_len 50, emb_dim 128 = 112 + 16, row-sharded over N GPUs).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      tools/bench_sharded.py [--items 50000000] [--seqs 4096] [--steps 20] [--warmup 5]

One step = one gather of S*T positions + one scatter-add of their gradients on every rank (the
K1+K3 / K13 pair of the hot path; the dense part of the step is data parallel and unchanged).  Prints
one JSON line on rank 0: positions/s over all ranks, per-GPU algorithmic GB/s of both kernels
(T*8 + 2*T*D*4 bytes per gathered window), timed with CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=50_000_000)
    ap.add_argument("--cates", type=int, default=9_400)
    ap.add_argument("--item-dim", type=int, default=112)
    ap.add_argument("--cate-dim", type=int, default=16)
    ap.add_argument("--seqs", type=int, default=4096)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args(argv)
    import torch
    import torch.distributed as dist
    from clsr_b200 import build, sharded as SH
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    Di, Dc, D = a.item_dim, a.cate_dim, a.item_dim + a.cate_dim
    item = SH.ShardedTable(a.items, Di, rank, world, local, dist=dist if world > 1 else None)
    cate = SH.ShardedTable(a.cates, Dc, rank, world, local, dist=dist if world > 1 else None)
    item.values.normal_(0, 0.01)
    cate.values.normal_(0, 0.01)
    g = np.random.default_rng(42 + rank)
    NB = 4
    feeds = []
    for _ in range(NB):   # zipf ids = popularity ranks, left-aligned windows padded with id 0
        ih = np.minimum(g.zipf(1.05, (a.seqs, a.T)), a.items - 1).astype(np.int32)
        ln = g.integers(1, a.T + 1, a.seqs)
        ih[np.arange(a.T)[None, :] >= ln[:, None]] = 0
        ch = (ih.astype(np.int64) * 2654435761 % (a.cates - 1) + 1).astype(np.int32)
        ch[ih == 0] = 0
        feeds.append((torch.from_numpy(ih).cuda(), torch.from_numpy(ch).cuda()))
    out = torch.empty(a.seqs, a.T, D, device="cuda")
    d = torch.randn(a.seqs, a.T, D, device="cuda")

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sync()
    times = {}
    for name, fn in (("gather", lambda f: SH.gather_history(item, cate, f[0], f[1], out=out)),
                     ("scatter_add", lambda f: SH.scatter_add_history(item, cate, f[0], f[1], d))):
        for i in range(a.warmup):
            fn(feeds[i % NB])
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            fn(feeds[(a.warmup + i) % NB])
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / a.steps
        if world > 1:
            t = torch.tensor([ms])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        times[name] = ms
    if rank == 0:
        pos = a.seqs * a.T
        alg = pos * (8 + 2 * D * 4)
        step_ms = times["gather"] + times["scatter_add"]
        print(json.dumps({
            "metric": "user-sequences/sec (row-sharded history gather + scatter-add only, seq_len=%d, emb_dim=%d)" % (a.T, D),
            "value": a.seqs * world / (step_ms / 1e3), "unit": "user-sequences/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "config 4: %d items x %d + %d cates x %d fp32, row r on rank r %% %d, %d sequences x %d "
                                   "positions per GPU, zipf ids; peer-memory gather (no all-to-all)"
                                   % (a.items, Di, a.cates, Dc, world, a.seqs, a.T),
                       "shard_bytes_per_gpu": item.local_rows * Di * 4 * 2},
            "kernels": {k: {"ms": v, "algorithmic_bytes": alg, "achieved_GBps_per_gpu": alg / 1e9 / (v / 1e3)}
                        for k, v in times.items()}}))
    if world > 1:
        dist.barrier()
    item.close(); cate.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
