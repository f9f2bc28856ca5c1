import sys, time, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
import fplll_b200 as fb
z = H.gold("r200_lll_update_gso.npz")
b = z["b"].copy()
t = time.time()
st, stats = fb.bkz_reduction(b, fb.BKZParam(60, strategies="default", flags=fb.BKZ_NO_LLL | fb.BKZ_MAX_LOOPS | fb.BKZ_VERBOSE, max_loops=int(sys.argv[1]) if len(sys.argv) > 1 else 1))
print("status", st, "wall", time.time() - t)
for k, v in stats.items():
    print("  ", k, v)
