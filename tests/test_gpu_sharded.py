"""Row-sharded tables (shard.cu) on the GPU: bit-exact against torch indexing / index_add_."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("Di,Dc,T", [(32, 8, 50), (112, 16, 50), (224, 32, 200)])
def test_world1_gather_bit_exact_and_scatter_add(cuda_lib, Di, Dc, T):
    """world = 1 runs the same kernels with one peer (itself): configs 1-5 row widths."""
    import torch
    from clsr_b200 import sharded as SH
    n_items, n_cates, S = 50021, 311, 64
    g = torch.Generator(device="cuda").manual_seed(Di + T)
    item = SH.ShardedTable(n_items, Di)
    cate = SH.ShardedTable(n_cates, Dc)
    item.values.copy_(torch.randn(n_items, Di, device="cuda", generator=g))
    cate.values.copy_(torch.randn(n_cates, Dc, device="cuda", generator=g))
    ih = torch.randint(0, n_items, (S, T), device="cuda", generator=g, dtype=torch.int32)
    ch = torch.randint(0, n_cates, (S, T), device="cuda", generator=g, dtype=torch.int32)
    ih[:, T // 2:] = 0                                      # padded tails: every row aims at id 0
    out = SH.gather_history(item, cate, ih, ch)
    ref = torch.cat([item.values[ih.long()], cate.values[ch.long()]], -1)
    assert torch.equal(out, ref)
    d = torch.randint(-3, 4, (S, T, Di + Dc), device="cuda", generator=g).float()   # integer-valued: exact sums
    item.zero_grad(); cate.zero_grad()
    SH.scatter_add_history(item, cate, ih, ch, d)
    torch.cuda.synchronize()
    gi = torch.zeros(n_items, Di, device="cuda").index_add_(0, ih.reshape(-1).long(), d.reshape(-1, Di + Dc)[:, :Di])
    gc = torch.zeros(n_cates, Dc, device="cuda").index_add_(0, ch.reshape(-1).long(), d.reshape(-1, Di + Dc)[:, Di:])
    assert torch.equal(item.grad, gi) and torch.equal(cate.grad, gc)
    item.close(); cate.close()


def test_bad_arguments_are_loud(cuda_lib):
    from clsr_b200 import sharded as SH
    from clsr_b200.engine import EngineError
    with pytest.raises(EngineError):
        SH.ShardedTable(100, 6)                             # rows must be 16-byte multiples
    with pytest.raises(EngineError):
        SH.ShardedTable(100, 8, rank=3, world=2)


def test_two_ranks_gather_and_scatter_through_peer_memory(cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SHARD_RESULT ")][-1]
    res = json.loads(line[len("SHARD_RESULT "):])
    assert res["ok"], res
