#!/bin/bash
# two GPUs of one box: hand-off between devices inside one process, shared ticket / radius push between two processes (CUDA IPC)
O=gpurun_out/r2
mkdir -p $O
nvidia-smi -L; nvidia-smi topo -m | head -6
echo "== in-process two-device test"
timeout 600 python -m pytest tests/test_enum_gpu.py -m gpu -q -p no:cacheprovider -k "two_devices or fixed_radius or bkz60_block" 2>&1 | cut -c1-300 | tail -15
echo "== single-block enumeration: 1 device, 2 devices in one process (hand-off), timing"
python - <<'PY'
import numpy as np, time, sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from fplll_b200 import enumeration as en
z = H.gold("enum_r200_b60_pruned_140.npz")
for devs in ([0], [0, 1]):
    for rep in range(3):
        t = time.perf_counter()
        r = en.enumerate_svp(z["mut"], z["rdiag"], z["pruning"], float(z["maxdist"]), devices=devs)
        dt = time.perf_counter() - t
    print("devices", devs, "sec %.4f" % dt, "nodes", int(r["nodes"].sum()), "equal ref", int(r["nodes"].sum()) == int(z["nodes"].sum()),
          "n_devices", r["stats"]["n_devices"], "dev_ms %.2f" % r["stats"]["device_ms"], "rounds", r["stats"]["n_rounds"])
PY
echo "== torchrun 2 ranks: bench.py (sharded enumeration over CUDA IPC + BKZ child on 2 devices)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err | cut -c1-300
python -c "
import json; j=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]); print('value', j['value'], 'enum', j.get('enum'), 'bkz60', {k:v for k,v in (j.get('bkz60') or {}).items() if k in ('wall_seconds','status','sec_enum','sec_lll_sizered','enum_nodes','attempts','error')})"
echo done
