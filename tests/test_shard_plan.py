"""Row-sharding plan of clsr_b200/sharded.py on CPU (gloo, world_size 2): ownership arithmetic, shard
extraction and the rank-major exchange of the per-rank IPC handle blobs.  A numpy emulation of the
fused peer-memory gather / scatter-add (what shard.cu computes) reproduces table[ids] and
np.add.at on the unsharded table."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clsr_b200 import sharded as SH  # noqa: E402


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n_rows", [1, 7, 64, 1001])
def test_plan_partitions_every_row_once(world, n_rows):
    ids = np.arange(n_rows)
    own, loc = SH.owner_of(ids, world), SH.local_row_of(ids, world)
    assert sum(SH.local_rows(n_rows, r, world) for r in range(world)) == n_rows
    table = np.arange(n_rows * 4, dtype=np.float32).reshape(n_rows, 4)
    shards = [SH.shard_of(table, r, world) for r in range(world)]
    for r in range(world):
        assert len(shards[r]) == SH.local_rows(n_rows, r, world)
    back = np.stack([shards[o][l] for o, l in zip(own, loc)])
    assert np.array_equal(back, table)


def test_emulated_peer_gather_and_scatter_match_unsharded():
    rng = np.random.default_rng(5)
    world, n_items, n_cates, Di, Dc, P = 4, 997, 33, 112, 16, 5000
    item = rng.standard_normal((n_items, Di)).astype(np.float32)
    cate = rng.standard_normal((n_cates, Dc)).astype(np.float32)
    ih = rng.integers(0, n_items, P).astype(np.int32)
    ch = rng.integers(0, n_cates, P).astype(np.int32)
    ish = [SH.shard_of(item, r, world) for r in range(world)]
    csh = [SH.shard_of(cate, r, world) for r in range(world)]
    got = np.concatenate([np.stack([ish[o][l] for o, l in zip(SH.owner_of(ih, world), SH.local_row_of(ih, world))]),
                          np.stack([csh[o][l] for o, l in zip(SH.owner_of(ch, world), SH.local_row_of(ch, world))])], 1)
    assert np.array_equal(got, np.concatenate([item[ih], cate[ch]], 1))
    d = rng.integers(-3, 4, (P, Di + Dc)).astype(np.float32)
    gi = [np.zeros_like(s) for s in ish]
    for o, l, row in zip(SH.owner_of(ih, world), SH.local_row_of(ih, world), d[:, :Di]):
        gi[o][l] += row
    ref = np.zeros_like(item)
    np.add.at(ref, ih, d[:, :Di])
    for r in range(world):
        assert np.array_equal(gi[r], SH.shard_of(ref, r, world))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * SH.HANDLE_BYTES
    allh = SH.exchange_handles(mine, world, dist)
    ok = len(allh) == world * SH.HANDLE_BYTES and all(
        allh[r * SH.HANDLE_BYTES:(r + 1) * SH.HANDLE_BYTES] == bytes([r + 1]) * SH.HANDLE_BYTES for r in range(world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_handle_exchange_is_rank_major_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        SH.ShardedTable(100, 16)
