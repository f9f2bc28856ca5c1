#!/bin/bash
O=gpurun_out/s2
mkdir -p $O
timeout 900 ncu --clock-control none --cache-control none -k regex:k_lll_cta --launch-skip 3000 --launch-count 1500 --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv --log-file $O/lll_list.csv python tools/gpurun_bkz_seed.py 40 1 > $O/ncu_lll_list.log 2>&1
tail -2 $O/ncu_lll_list.log | cut -c1-200; wc -l $O/lll_list.csv
