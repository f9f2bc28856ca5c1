#!/bin/bash
# Round-2 probe 4: library-free rounding / scaling, conflict-free mu cache, chunked serial chains; the MatGSO shim.
O=gpurun_out/r2
mkdir -p $O
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -40 > $O/t_all_4.log; tail -25 $O/t_all_4.log
echo "== BKZ-60"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v4.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other\|sec_get\|sec_ops" $O/bkz60_v4.txt
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof4.txt 2>&1
grep -A3 "LLL profile" $O/bkz60_prof4.txt | head -8; grep "wall\|sec_lll" $O/bkz60_prof4.txt
echo "== bench (no bkz)"
timeout 300 python bench.py --no-bkz --no-cpu-baseline > $O/bench_v4.json 2> $O/bench_v4.err; python -c "
import json; j=json.loads(open('$O/bench_v4.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e']['value'], j.get('householder',{}).get('frac_of_hbm_peak'))"
echo done
