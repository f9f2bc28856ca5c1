import sys, time, os, subprocess, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H
from fplll_b200 import enumeration as en
from test_enum_oracle import gso_block
names = sys.argv[1:]
for name in names:
    z = H.gold(name)
    for rep in range(2):
        t = time.perf_counter(); res = en.enumerate_svp(z["mut"], z["rdiag"], z["pruning"], float(z["maxdist"])); dt = time.perf_counter() - t
    st = res["stats"]; n = int(res["nodes"].sum())
    print("%s bpsm=%s b0=%s mul=%s: wall %.4fs dev %.2fms host %.0fus nodes %.3g rate %.3g/s rounds %d roots %d" % (name[:24], os.environ.get("B200_ENUM_BLOCKS_PER_SM"), os.environ.get("B200_ENUM_BUDGET0"), os.environ.get("B200_ENUM_BUDGET_MUL"), dt, st["device_ms"], st["host_breadth_us"], n, n / dt, st["n_rounds"], st["n_roots"]), flush=True)
# small blocks typical of BKZ preprocessing: block [100,140) of the r200 basis with beta=40 default pruning at 1.1 GH
