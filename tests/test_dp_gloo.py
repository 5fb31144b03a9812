"""Data-parallel sharding plan on CPU (gloo, world_size 2): the exchanges the multi-GPU engine
performs -- BatchNorm sums in forward (and, through autograd, backward), the contrastive row count,
summed gradients -- reproduce the single-process result on the concatenated batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.distributed.nn.functional as dfn
    import parity_util as PU
    from oracle import clsr_oracle as O
    torch.set_num_threads(2)
    G, S = 5, 8
    feed, prm = PU.small_problem(S=S, G=G, T=12, seed=13, n_items=300, n_cates=12, n_users=40)
    cfg = PU.oracle_config(G, max_seq_length=12, contrastive_length_threshold=3)
    rows = slice(rank * (S // world) * G, (rank + 1) * (S // world) * G)
    shard = {k: v[rows] for k, v in feed.items()}

    class Ctx:
        pass
    ctx = Ctx()
    ctx.world = world
    ctx.all_reduce = lambda t: dfn.all_reduce(t, op=dist.ReduceOp.SUM)

    def grads(batch, dp):
        p = {k: torch.as_tensor(v).double() for k, v in prm.items()}
        for k, t in p.items():
            if "moving_" not in k:
                t.requires_grad_(True)
        if dp:
            with O.data_parallel(ctx):
                out = O.forward(p, batch, cfg, True, torch.float64)
                L = O.losses(out, p, batch, cfg, torch.float64)
        else:
            out = O.forward(p, batch, cfg, True, torch.float64)
            L = O.losses(out, p, batch, cfg, torch.float64)
        (L["data_loss"] + L["contrastive_loss"]).backward()
        g = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in p.items() if t.requires_grad}
        return L, g, out

    L, g, out = grads(shard, True)
    for t in g.values():
        dist.all_reduce(t)
    data = L["data_loss"].detach().clone()
    dist.all_reduce(data)
    con = L["contrastive_loss"].detach().clone()   # local partial sum over the global row count
    dist.all_reduce(con)
    Lf, gf, outf = grads(feed, False)
    worst = max(float((g[k] - gf[k]).abs().max() / max(float(gf[k].abs().max()), 1e-12)) for k in gf
                if float(gf[k].abs().max()) > 1e-12)
    lerr = abs(float(data) - float(Lf["data_loss"]))
    cerr = abs(float(con) - float(Lf["contrastive_loss"]))
    # the shard's logits equal the corresponding rows of the full-batch run (sync-BN forward)
    ferr = float((out["logit"].detach() - outf["logit"].detach()[rows]).abs().max())
    result[rank] = (worst, lerr, cerr, ferr)
    dist.destroy_process_group()


def _spawn_target(rank, world, port, q):
    res = {}
    _worker(rank, world, port, res)
    q.put((rank, res[rank]))


def test_sharded_step_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_spawn_target, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, (worst, lerr, cerr, ferr) in got:
        assert worst < 1e-8 and lerr < 1e-10 and cerr < 1e-10 and ferr < 1e-10, (rank, worst, lerr, cerr, ferr)
