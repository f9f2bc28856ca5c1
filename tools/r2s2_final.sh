#!/bin/bash
# final single-GPU pass of the round: GPU suite, smoke, bench line, ncu launch list of the bench command
O=gpurun_out/s2
mkdir -p $O
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
echo "== bench"
timeout 1200 python bench.py > $O/bench_final.json 2> $O/bench_final.err; tail -2 $O/bench_final.err | cut -c1-300
python - <<'PY'
import json
j = json.loads(open('gpurun_out/s2/bench_final.json').read().strip().splitlines()[-1])
print('value', j['value'], 'roofline', j['roofline'])
print('e2e', j['e2e'], 'clocks', j['clocks'])
print('enum', {k: v for k, v in (j.get('enum') or {}).items() if k != 'workload'})
print('hh', {k: v for k, v in (j.get('householder') or {}).items() if k != 'workload'})
print('bkz60', {k: v for k, v in (j.get('bkz60') or {}).items() if k != 'workload'})
print('cpu', {k: v for k, v in (j.get('cpu_baseline') or {}).items() if k != 'sample'})
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 > $O/bench_ref.json 2>/dev/null; cut -c1-400 $O/bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 4 --warmup 1 --no-extras --no-cpu-baseline --no-traffic > $O/ncu_launch.log 2>&1
grep -c "k_update_row\|k_row_op_end" $O/r2_launches_bench.csv
echo done
