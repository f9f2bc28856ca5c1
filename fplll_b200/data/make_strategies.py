"""Converts the reference's BKZ strategy table (strategies/default.json: for every block size the preprocessing block
sizes and 21 pruning vectors computed offline by fplll's pruner) into strategies_default.npz, the array form
b200bkz_add_strategy takes.  Run once in the build container:  python fplll_b200/data/make_strategies.py
Data only — the pruner that produced it is out of scope."""
import json
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/strategies/default.json"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "strategies_default.npz")
pack = {}
sizes = []
for e in json.load(open(src)):
    bs = int(e["block_size"])
    sizes.append(bs)
    pack["pre_%d" % bs] = np.array(e.get("preprocessing_block_sizes", []), dtype=np.int32)
    pp = e.get("pruning_parameters", [])
    pack["ghf_%d" % bs] = np.array([p[0] for p in pp], dtype=np.float64)
    pack["exp_%d" % bs] = np.array([p[2] for p in pp], dtype=np.float64)
    pack["coef_%d" % bs] = np.array([p[1] for p in pp], dtype=np.float64).reshape(len(pp), bs if pp else 0)
pack["block_sizes"] = np.array(sizes, dtype=np.int32)
np.savez_compressed(out, **pack)
print(out, os.path.getsize(out), "bytes;", len(sizes), "block sizes")
