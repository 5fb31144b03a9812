"""Developer tool: per-buffer relative error of the CUDA step against the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import parity_util as PU


def main():
    G = int(os.environ.get("G", 5))
    group = int(os.environ.get("GROUP", G))
    S = int(os.environ.get("S", 24))
    feed, prm = PU.small_problem(S=S, G=G, seed=int(os.environ.get('SEED', 3)))
    feed = PU.set_lengths(feed, [1, 50, 3, 5, 6, 2], G)
    eng = PU.make_engine(prm, 3000, 40, 200, max_rows=S * G, G=G, math_mode=int(os.environ.get('MATH', 0)))
    eng.set_debug_sync(True)
    res, losses = PU.compare_step(eng, feed, prm, G, group, metric=PU.relerr_l2 if os.environ.get('L2') else None)
    print("losses", losses)
    bad = 0
    for k, v in res.items():
        flag = "" if v < 2e-3 else "   <<<<<<"
        bad += v >= 2e-3
        print("%-70s %.3e%s" % (k, v, flag))
    print("BAD", bad, "launches", eng.kernel_launches())


if __name__ == "__main__":
    main()
