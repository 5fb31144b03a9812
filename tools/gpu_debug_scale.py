"""Developer aid: per-intermediate deviation of the CUDA step from the oracle at growing batch sizes.
  python tools/gpu_debug_scale.py [S ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as PU  # noqa: E402
from clsr_b200 import params as P, synth  # noqa: E402

NI, NC, NU = 4_000_000, 9_400, 1_000_000
sizes = [int(x) for x in sys.argv[1:]] or [256, 1024, 4096]
prm = PU.scale_params(P.init_params(NI, NC, NU, seed=101), 101)
for S in sizes:
    src = synth.SyntheticSource(NI, NC, NU, 50, seed=101)
    feed = src.batch(S, 4)
    for mode in (0, 1):
        for dt in ((torch.float64, torch.float32) if S <= 1024 else (torch.float32,)):
            eng = PU.make_engine(prm, NI, NC, NU, max_rows=S * 5, G=5, math_mode=mode)
            t0 = time.time()
            res, _ = PU.compare_step(eng, feed, prm, 5, 5, metric=PU.relerr_l2, dtype=dt)
            worst = sorted(((v, k) for k, v in res.items() if not k.endswith("b_nn_output")), reverse=True)[:8]
            print("S=%d mode=%d oracle=%s (%.0fs):" % (S, mode, str(dt)[-7:], time.time() - t0),
                  "  ".join("%s=%.1e" % (k, v) for v, k in worst), flush=True)
            print("   fwd:", "  ".join("%s=%.1e" % (k[4:], v) for k, v in res.items() if k.startswith("fwd/")), flush=True)
            eng.close()
