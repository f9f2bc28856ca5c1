#!/bin/bash
O=gpurun_out/r2
mkdir -p $O
echo "== ncu dense sampling of one long k_lll_cta launch"
timeout 900 ncu --cache-control none --clock-control none -k regex:k_lll_cta --launch-skip 4002 --launch-count 1 \
  --section SourceCounters --warp-sampling-interval 1 -f -o /tmp/lll_src python tools/gpurun_bkz_seed.py 60 1 > $O/ncu_lll10.log 2>&1
tail -2 $O/ncu_lll10.log
ncu -i /tmp/lll_src.ncu-rep --page source --csv 2>/dev/null > /tmp/lll_sass.csv; wc -l /tmp/lll_sass.csv
python tools/ncu_sass_samples.py < /tmp/lll_sass.csv > $O/ncu_lll_sass10.txt 2>&1; head -5 $O/ncu_lll_sass10.txt | cut -c1-200
ncu -i /tmp/lll_src.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for i,c in enumerate(h):
        if 'gpu__time_duration' in c or 'launch__grid' in c: print(c, r[i])
"
echo done
