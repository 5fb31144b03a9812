"""Worker of the multi-GPU parity test: launched with torch.distributed.run, one rank per GPU.
Each rank trains on its shard of the sequences; rank 0 also trains a single-GPU engine on the whole
batch and compares losses and final variables."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as PU  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    G, S, steps = 5, 16 * world, 3
    NI, NC, NU = 3000, 40, 200
    feeds, prm = [], None
    for i in range(steps):
        f, p = PU.small_problem(S=S, G=G, seed=31 + i)
        feeds.append(f)
        prm = prm or p
    from clsr_b200.engine import Engine
    per = S // world
    kw = dict(seq_len=50, train_group=G, optimizer=os.environ.get("OPT", "adam"), device=local)
    eng = Engine(NI, NC, NU, max_rows=per * G, **kw)
    eng.set_params(prm)
    shard = os.environ.get("SHARD", "1") == "1"   # row-sharded tables + peer-memory reductions, or replicated tables
    eng.comm_init(rank, world, dist, shard=shard)
    assert eng.sharded == shard
    rows = slice(rank * per * G, (rank + 1) * per * G)
    losses = [eng.train_step({k: v[rows] for k, v in f.items()}, group=G) for f in feeds]
    mine = eng.get_params()
    ok, report = True, {}
    if rank == 0:
        ref = Engine(NI, NC, NU, max_rows=S * G, **kw)
        ref.set_params(prm)
        want = [ref.train_step(f, group=G) for f in feeds]
        full = ref.get_params()
        for a, b in zip(losses, want):
            for k in a:
                if abs(a[k] - b[k]) > 2e-5 * max(abs(b[k]), 1e-3):
                    ok = False
                    report["loss_" + k] = (a[k], b[k])
        for k, v in full.items():
            d0 = (v.reshape(-1) - np.asarray(prm[k], np.float32).reshape(-1)).astype(np.float64)
            d1 = (mine[k].reshape(-1) - v.reshape(-1)).astype(np.float64)
            moved, err = np.abs(d0).max(), np.abs(d1).max()
            if k.endswith("att_fcn/nn_part/b_nn_output") or k.endswith("logit_fcn/nn_part/b_nn_output"):
                continue
            # Adam moves an element whose gradient is rounding noise by +-lr per step whatever the noise is, so single
            # elements may differ by a few lr between two summation orders: the variable as a whole must agree
            # (relative L2 of the update), and no element may be off by more than the largest update
            if np.linalg.norm(d1) > 0.03 * np.linalg.norm(d0) + 1e-7 or err > 0.7 * moved + 1e-7:
                ok = False
                report[k] = (float(err), float(moved), float(np.linalg.norm(d1)), float(np.linalg.norm(d0)))
    # all replicas must hold identical variables
    flat = torch.from_numpy(np.concatenate([mine[k].reshape(-1) for k in sorted(mine)]).astype(np.float64))
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    if rank == 0:
        print("DP_RESULT " + json.dumps({"ok": ok and spread < 1e-5, "replica_spread": spread, "report": report,
                                         "last_loss": losses[-1], "sharded": shard, "world": world}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
