#!/bin/bash
# session-2 status snapshot on one B200: bench line (with BKZ-60 + reference), then the GPU suite
O=gpurun_out/s2
mkdir -p $O
nvidia-smi -L
echo "== bench (N=1, default)"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -3 $O/bench_n1.err | cut -c1-300
python - <<'PY'
import json
j = json.loads(open('gpurun_out/s2/bench_n1.json').read().strip().splitlines()[-1])
print('value', j['value'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'])
print('enum', j.get('enum'))
print('hh', j.get('householder'))
print('bkz60', j.get('bkz60'))
print('cpu', j.get('cpu_baseline'))
PY
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=15 2>&1 | tail -30
echo done
