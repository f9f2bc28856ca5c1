#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof8.txt 2>&1
grep -A5 "LLL profile" $O/bkz60_prof8.txt | head -6; grep "wall\|sec_lll" $O/bkz60_prof8.txt
echo done
