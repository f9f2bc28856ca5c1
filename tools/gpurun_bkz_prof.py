import sys, time, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
import fplll_b200 as fb
z = H.gold("r200_lll_update_gso.npz")
bs = int(sys.argv[1]); seed = int(sys.argv[2])
b = z["b"].copy()
t = time.time()
st, stats = fb.bkz_reduction(b, fb.BKZParam(bs, strategies="default", flags=fb.BKZ_NO_LLL | fb.BKZ_MAX_LOOPS, max_loops=1, seed=seed))
print("BKZ-%d seed %d status %d wall %.1f" % (bs, seed, st, time.time() - t))
for k, v in stats.items():
    print("   ", k, v)
