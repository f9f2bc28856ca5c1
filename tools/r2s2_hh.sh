#!/bin/bash
O=gpurun_out/s2
mkdir -p $O
echo "== HH parity"
timeout 900 python -m pytest tests/test_hh_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
echo "== bench (no BKZ): householder + enum entries"
timeout 600 python bench.py --no-bkz --no-cpu-baseline --steps 50 > $O/bench_hh.json 2> $O/bench_hh.err; tail -2 $O/bench_hh.err | cut -c1-300
python - <<'PY'
import json
j = json.loads(open('gpurun_out/s2/bench_hh.json').read().strip().splitlines()[-1])
print('hh', j.get('householder'))
print('enum', {k: v for k, v in (j.get('enum') or {}).items() if k != 'workload'})
PY
echo "== old kernel for comparison"
B200_HH_X32=0 timeout 600 python bench.py --no-bkz --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'): print('hh', json.loads(ln).get('householder'))"
echo done
