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
if os.environ.get("SWEEP_SMALL"):
    from fplll_b200.bkz import load_strategies
    from oracle import oracle as O
    b = H.gold("r200_lll_update_gso.npz")["b"]
    strat = load_strategies()
    m = O.OracleGSO(b, 0); assert m.update_gso(); s = m.state()
    for bs in (20, 30, 40, 50):
        tot_t, tot_n, calls, rounds, dev, hb, tu = 0.0, 0, 0, 0, 0.0, 0.0, 0.0
        coef = strat[bs][3]
        prun = coef[0] if len(coef) else None
        for first in range(60, 140, 8):
            mut = np.zeros((bs, bs))
            for k in range(bs):
                for j in range(k + 1, bs):
                    mut[k, j] = s["mu"][first + j, first + k]
            rdiag = np.array([s["r"][first + i, first + i] for i in range(bs)])
            t = time.perf_counter(); res = en.enumerate_svp(mut, rdiag, prun, float(rdiag[0] * 0.99)); dt = time.perf_counter() - t
            tot_t += dt; tot_n += int(res["nodes"].sum()); calls += 1; rounds += res["stats"]["n_rounds"]; dev += res["stats"]["device_ms"]; hb += res["stats"]["host_breadth_us"]; tu += res["stats"]["total_us"]
        print("small bs=%d yield=%s: %.1f us/call, %.0f nodes/call, %.1f rounds/call, dev %.1f us, host breadth %.1f us, C total %.1f us" % (bs, os.environ.get("B200_ENUM_YIELD"), 1e6 * tot_t / calls, tot_n / calls, rounds / calls, 1e3 * dev / calls, hb / calls, tu / calls), flush=True)
