#!/bin/bash
O=gpurun_out/s2
mkdir -p $O
timeout 900 python bench.py > $O/bench_final2.json 2> $O/bench_final2.err; tail -2 $O/bench_final2.err | cut -c1-300
python - <<'PY'
import json
j = json.loads(open('gpurun_out/s2/bench_final2.json').read().strip().splitlines()[-1])
print('value', j['value'], 'frac', j['roofline']['frac'], 'traffic', j['roofline']['traffic'], 'e2e', j['e2e']['value'], 'clocks', j['clocks'])
print('enum', {k: v for k, v in (j.get('enum') or {}).items() if k not in ('workload','cpu_reference')})
print('hh', {k: v for k, v in (j.get('householder') or {}).items() if k != 'workload'})
print('bkz60', {k: v for k, v in (j.get('bkz60') or {}).items() if k != 'workload'})
PY
