#!/bin/bash
for L in lib lib_ps32 lib_ps100; do
echo "== $L"
B200_LIB_DIR=$L timeout 200 python tools/gpurun_bkz_prof.py 60 2 2>&1 | grep "BKZ-60\|sec_lll\|sec_enum\|sec_total\|sec_other"
done
echo done
