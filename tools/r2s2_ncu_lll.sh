#!/bin/bash
# source-level profile of ONE long launch of the single-lattice LLL kernel inside a BKZ-60 tour
O=gpurun_out/s2
mkdir -p $O
timeout 900 ncu --cache-control none --clock-control none -k regex:k_lll_cta --launch-skip ${SKIP:-6000} --launch-count 1 \
  --section SourceCounters --section WarpStateStats --section LaunchStats --import-source on -f -o $O/lll_src python tools/gpurun_bkz_seed.py 60 1 > $O/ncu_lll.log 2>&1
tail -3 $O/ncu_lll.log | cut -c1-200
ls -la $O/lll_src.ncu-rep
