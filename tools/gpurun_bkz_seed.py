import sys, time, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
import fplll_b200 as fb
z = H.gold("r200_lll_update_gso.npz")
for seed in [int(s) for s in sys.argv[2:]]:
    b = z["b"].copy()
    t = time.time()
    st, stats = fb.bkz_reduction(b, fb.BKZParam(60, strategies="default", flags=fb.BKZ_NO_LLL | fb.BKZ_MAX_LOOPS | 8, max_loops=1, max_time=float(sys.argv[1]), seed=seed))
    print("seed", seed, "status", st, "wall %.1f" % (time.time() - t), "enum_calls", stats["enum_calls"], "nodes %.3g" % stats["enum_nodes"], "lll %.1f enum %.1f other %.1f" % (stats["sec_lll"], stats["sec_enum"], stats["sec_other"]), flush=True)
