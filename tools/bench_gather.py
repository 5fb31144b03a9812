"""Standalone history gather / scatter-add (K1+K3, K13) at a size where launch ramp does not matter.
  python tools/bench_gather.py            (CLSR_GATHER_LDG=1 selects the register-path gather for A/B)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clsr_b200 import synth  # noqa: E402
from clsr_b200.engine import Engine, TABLE_CATE, TABLE_ITEM  # noqa: E402

NI, NC, T = 4_000_000, 9_400, 50
npos = 32 * 4096 * T
eng = Engine(NI, NC, 1000, max_rows=4096, seq_len=T, training=False)
g = torch.Generator(device="cuda").manual_seed(0)
eng.tables[TABLE_ITEM].normal_(generator=g)
eng.tables[TABLE_CATE].normal_(generator=g)
src = synth.SyntheticSource(NI, NC, 1000, T, seed=3)
res = {}
for kind in ("zipf", "uniform"):
    if kind == "zipf":
        ih = torch.from_numpy(np.concatenate([src.lines(4096)["item_history"] for _ in range(32)])).cuda().reshape(-1)
        ch = torch.from_numpy(src.cate_of(ih.cpu().numpy())).cuda()
    else:
        ih = torch.randint(1, NI, (npos,), device="cuda", dtype=torch.int32, generator=g)
        ch = torch.randint(1, NC, (npos,), device="cuda", dtype=torch.int32, generator=g)
    out = torch.empty(npos, 40, device="cuda")
    call = lambda: eng._check(eng.lib.clsr_gather_history(eng.h, ih.data_ptr(), ch.data_ptr(), npos, out.data_ptr()))
    call()
    eng.synchronize()
    ref = torch.cat([eng.tables[TABLE_ITEM][ih.long()], eng.tables[TABLE_CATE][ch.long()]], -1)
    assert torch.equal(out, ref), "gather not bit-exact"
    del ref
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        call()
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = npos * (8 + 2 * 40 * 4)
    res["gather_" + kind] = {"ms": round(ms, 4), "GBps": round(nbytes / 1e9 / (ms / 1e3), 1), "positions": npos}
    # scatter-add of a [positions, 40] gradient into compact unique rows (clsr_scatter_history_grad: unique + zero + scatter)
    if npos <= 4096 * T * 4 or True:
        n2 = 4096 * T   # one step's worth (the compact buffers are sized for one batch)
        d = torch.randn(n2, 40, device="cuda", generator=g)
        ih2, ch2 = ih[:n2].contiguous(), ch[:n2].contiguous()
        sc = lambda: eng._check(eng.lib.clsr_scatter_history_grad(eng.h, ih2.data_ptr(), ch2.data_ptr(), n2, d.data_ptr()))
        eng.set_profiling(True)
        for _ in range(13):
            sc()
        prof = eng.profile()
        eng.set_profiling(False)
        ms2 = prof["scatter_hist"][0] / prof["scatter_hist"][1]
        nb2 = n2 * (8 + 2 * 40 * 4)
        res["scatter_" + kind] = {"ms": round(ms2, 4), "GBps": round(nb2 / 1e9 / (ms2 / 1e3), 1), "positions": n2}
print(json.dumps({"ldg": bool(os.environ.get("CLSR_GATHER_LDG")), **res}))
