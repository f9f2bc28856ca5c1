#!/bin/bash
# Round-2 probe 3: is the single-lattice LLL kernel instruction-fetch bound (138 k SASS instructions)?  ncu warp-state
# capture without cache flushing, the noinline build, enumerator round tuning.
O=gpurun_out/r2
mkdir -p $O
echo "== gpu tests (new direct tests)"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 > $O/t_all_3.log; tail -4 $O/t_all_3.log
echo "== BKZ-60: default vs noinline build"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v3.txt 2>&1; grep "wall\|sec_lll\|sec_enum\|sec_other" $O/bkz60_v3.txt
B200_LIB_DIR=lib_ni timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_v3_noinline.txt 2>&1; grep "wall\|sec_lll\|sec_enum" $O/bkz60_v3_noinline.txt
echo "== enumerator tuning (sec_enum of a BKZ-60 tour)"
for ys in 8 16 32; do for mr in 128 256; do
  B200_ENUM_YIELD_SMALL=$ys B200_ENUM_MIN_ROOTS=$mr timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_ys${ys}_mr${mr}.txt 2>&1
  echo "yield_small=$ys min_roots=$mr: $(grep 'sec_enum\|^status' $O/bkz60_ys${ys}_mr${mr}.txt | tr '\n' ' ')"
done; done
echo "== ncu warp states of k_lll_cta, caches not flushed"
timeout 600 ncu --cache-control none --clock-control none -k regex:k_lll_cta --launch-skip 4000 --launch-count 6 \
  --section WarpStateStats --section SchedulerStats --section InstructionStats --section LaunchStats --section SourceCounters \
  --import-source on -f -o $O/lll_cta_r2 python tools/gpurun_bkz_seed.py 60 1 > $O/ncu_lll.log 2>&1
tail -3 $O/ncu_lll.log
ncu -i $O/lll_cta_r2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
keep=[i for i,h in enumerate(hdr) if any(k in h for k in ('Kernel Name','gpu__time_duration','issue_stalled','smsp__warps_issue','smsp__inst_executed.sum','smsp__cycles_active.avg','inst_issued'))]
for r in rows[:1]+rows[2:]:
    print(' | '.join(hdr[i]+'='+r[i] for i in keep if 'stalled' not in hdr[i])[:600])
    st=sorted(((float(r[i].replace(',','')) if r[i].replace(',','').replace('.','').isdigit() else 0.0, hdr[i]) for i in keep if 'issue_stalled' in hdr[i] and 'ratio' in hdr[i]), reverse=True)[:8]
    print('   top stalls:', [(round(v,2), h.split('issue_stalled_')[1].split('_per')[0] if 'issue_stalled_' in h else h) for v,h in st])
" | tee $O/ncu_lll_summary.txt | head -40
echo done
