#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
echo "== shim + bkz60 tests"
timeout 2400 python -m pytest tests/test_shim_gpu.py "tests/test_bkz_gpu.py::test_bkz60_default_strategies_on_dim200_one_tour_quality" -m gpu -q -p no:cacheprovider --durations=8 2>&1 | cut -c1-400 > $O/t_shim_5.log; tail -70 $O/t_shim_5.log
echo "== ncu source-level samples of one long k_lll_cta launch"
timeout 900 ncu --cache-control none --clock-control none -k regex:k_lll_cta --launch-skip 4002 --launch-count 1 \
  --section SourceCounters --section WarpStateStats --import-source on -f -o /tmp/lll_src python tools/gpurun_bkz_seed.py 60 1 > $O/ncu_lll5.log 2>&1
tail -2 $O/ncu_lll5.log; ls -la /tmp/lll_src.ncu-rep
ncu -i /tmp/lll_src.ncu-rep --page source --csv --print-source cuda 2>/dev/null > /tmp/lll_src_cuda.csv; wc -l /tmp/lll_src_cuda.csv; head -c 1500 /tmp/lll_src_cuda.csv
python tools/ncu_lines.py 90 < /tmp/lll_src_cuda.csv > $O/ncu_lll_lines.txt 2>&1; head -100 $O/ncu_lll_lines.txt
ncu -i /tmp/lll_src.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for i,c in enumerate(h):
        if 'gpu__time_duration' in c or 'inst_issued' in c or 'issue_stalled' in c and 'ratio' in c: print(c, r[i])
" | head -40
echo done
