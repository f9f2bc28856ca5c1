#!/bin/bash
# Round-2 probe 6: streamed wavefronts (values published per step, no barrier per panel) in the CTA LLL.
O=gpurun_out/r2
mkdir -p $O
echo "== gpu tests (LLL / BKZ / shim quick)"
timeout 1500 python -m pytest tests/test_gso_gpu.py tests/test_bkz_gpu.py tests/test_shim_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | cut -c1-300 | tail -30 > $O/t_6.log; tail -12 $O/t_6.log
echo "== BKZ-60"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v6.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v6.txt
B200_LIB_DIR=lib_prof timeout 400 python tools/gpurun_bkz60_trial.py > $O/bkz60_prof6.txt 2>&1
grep -A3 "LLL profile" $O/bkz60_prof6.txt | head -8; grep "wall\|sec_lll" $O/bkz60_prof6.txt
echo "== ncu SASS-level samples of one long k_lll_cta launch"
timeout 900 ncu --cache-control none --clock-control none -k regex:k_lll_cta --launch-skip 4002 --launch-count 1 \
  --section SourceCounters --section WarpStateStats -f -o /tmp/lll_src python tools/gpurun_bkz_seed.py 60 1 > $O/ncu_lll6.log 2>&1
tail -2 $O/ncu_lll6.log
ncu -i /tmp/lll_src.ncu-rep --page source --csv 2>/dev/null > /tmp/lll_sass.csv; wc -l /tmp/lll_sass.csv
python tools/ncu_sass_samples.py < /tmp/lll_sass.csv > $O/ncu_lll_sass6.txt 2>&1; head -30 $O/ncu_lll_sass6.txt | cut -c1-250
cuobjdump -lelf fplll_b200/lib/libb200bkz.so | head -3
echo done
