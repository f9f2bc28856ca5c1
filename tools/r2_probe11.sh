#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
echo "== BKZ-60: noinline operations build"
B200_LIB_DIR=lib_ni timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v11_ni.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v11_ni.txt
echo "== BKZ-60: default build"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v11.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v11.txt
echo done
