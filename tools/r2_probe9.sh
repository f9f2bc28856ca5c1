#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
echo "== gpu tests (LLL / BKZ)"
timeout 1500 python -m pytest tests/test_gso_gpu.py tests/test_bkz_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | cut -c1-300 | tail -30 > $O/t_9.log; tail -6 $O/t_9.log
echo "== BKZ-60"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v9.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v9.txt
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof9.txt 2>&1
grep -A5 "LLL profile" $O/bkz60_prof9.txt | head -6; grep "wall\|sec_lll" $O/bkz60_prof9.txt
echo done
