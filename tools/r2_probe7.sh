#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
echo "== gpu tests (LLL / BKZ / HH)"
timeout 1500 python -m pytest tests/test_gso_gpu.py tests/test_bkz_gpu.py tests/test_hh_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | cut -c1-300 | tail -30 > $O/t_7.log; tail -8 $O/t_7.log
echo "== BKZ-60"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v7.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v7.txt
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof7.txt 2>&1
grep -A3 "LLL profile" $O/bkz60_prof7.txt | head -4; grep "wall\|sec_lll" $O/bkz60_prof7.txt
echo "== bench (no bkz): householder with chunked chain_sum"
timeout 400 python bench.py --no-bkz --no-cpu-baseline > $O/bench_v7.json 2> $O/bench_v7.err; python -c "
import json; j=json.loads(open('$O/bench_v7.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e']['value'], j.get('householder'))"
echo done
