"""Worker of the row-sharded table test: launched with torch.distributed.run, one rank per GPU.
Every rank gathers its own ids through the peers' shards and scatter-adds integer-valued gradients
into the owners' gradient shards; results are compared bit-exactly with the unsharded table."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    from clsr_b200 import sharded as SH
    n_items, n_cates, Di, Dc, S, T = 200003, 977, 112, 16, 512, 50      # config-4 row widths
    rng = np.random.default_rng(11)
    item_full = rng.standard_normal((n_items, Di)).astype(np.float32)
    cate_full = rng.standard_normal((n_cates, Dc)).astype(np.float32)
    item = SH.ShardedTable(n_items, Di, rank, world, local, dist=dist)
    cate = SH.ShardedTable(n_cates, Dc, rank, world, local, dist=dist)
    item.load_global(item_full)
    cate.load_global(cate_full)
    torch.cuda.synchronize()
    dist.barrier()                                   # owners' shards are in place before any peer reads
    feeds = []
    for r in range(world):                           # every rank can rebuild every rank's feed
        g = np.random.default_rng(100 + r)
        ih = np.minimum(g.zipf(1.3, (S, T)) - 1, n_items - 1).astype(np.int32)
        ch = g.integers(0, n_cates, (S, T)).astype(np.int32)
        d = g.integers(-4, 5, (S, T, Di + Dc)).astype(np.float32)
        feeds.append((ih, ch, d))
    ih, ch, d = (torch.from_numpy(x).cuda() for x in feeds[rank])
    out = SH.gather_history(item, cate, ih, ch)
    torch.cuda.synchronize()
    ref = np.concatenate([item_full[feeds[rank][0]], cate_full[feeds[rank][1]]], -1)
    ok_gather = bool(np.array_equal(out.cpu().numpy(), ref))
    item.zero_grad(); cate.zero_grad()
    torch.cuda.synchronize()
    dist.barrier()                                   # gradient shards are zero before any peer adds
    SH.scatter_add_history(item, cate, ih, ch, d)
    torch.cuda.synchronize()
    dist.barrier()                                   # every peer's reductions have landed
    gi = np.zeros_like(item_full)
    gc = np.zeros_like(cate_full)
    for fih, fch, fd in feeds:
        np.add.at(gi, fih.reshape(-1), fd.reshape(-1, Di + Dc)[:, :Di])
        np.add.at(gc, fch.reshape(-1), fd.reshape(-1, Di + Dc)[:, Di:])
    ok_scatter = bool(np.array_equal(item.grad.cpu().numpy()[:len(gi[rank::world])], gi[rank::world]) and
                      np.array_equal(cate.grad.cpu().numpy()[:len(gc[rank::world])], gc[rank::world]))
    flags = [None] * world
    dist.all_gather_object(flags, (ok_gather, ok_scatter))
    if rank == 0:
        print("SHARD_RESULT " + json.dumps({"ok": all(a and b for a, b in flags), "per_rank": flags, "world": world}))
    dist.barrier()
    item.close(); cate.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
