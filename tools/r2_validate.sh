#!/bin/bash
# Round-2 validation queue: everything written without GPU time at the end of round 1, in one gpurun call.
#   gpurun --timeout 2400 -- 'bash tools/r2_validate.sh'
O=gpurun_out/r2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
lscpu | egrep 'Model name|^CPU\(s\)|NUMA node\(s\)' >> $O/gpus.txt

echo "== 1. full -m gpu suite, experimental tests un-gated"
B200_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > $O/t_all_exp.log
tail -5 $O/t_all_exp.log

echo "== 2. variant build (-DB200_MU_CACHE=1 -DB200_CTA_MOVE=1): LLL / BKZ parity, cache off and on"
B200_LIB_DIR=lib_exp timeout 900 python -m pytest tests/test_gso_gpu.py tests/test_bkz_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > $O/t_exp_move.log
tail -3 $O/t_exp_move.log
B200_LIB_DIR=lib_exp B200_LLL_MU_SMEM=1 timeout 900 python -m pytest tests/test_gso_gpu.py tests/test_bkz_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > $O/t_exp_mucache.log
tail -3 $O/t_exp_mucache.log

echo "== 3. BKZ-60 one tour: default / CTA move / CTA move + mu cache"
timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_default.txt 2>&1
B200_LIB_DIR=lib_exp timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_move.txt 2>&1
B200_LIB_DIR=lib_exp B200_LLL_MU_SMEM=1 timeout 300 python tools/gpurun_bkz60_trial.py > $O/bkz60_move_mucache.txt 2>&1
grep -H "wall\|sec_lll\|sec_enum\|status" $O/bkz60_*.txt

echo "== 4. microbench incl. blocked Gram; update_R cta32; TMA update kernel"
B200_TEST_EXPERIMENTAL=1 timeout 600 python tools/microbench.py > $O/microbench_exp.txt 2>&1
grep "M2" $O/microbench_exp.txt
timeout 300 python bench.py --no-bkz --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
B200_HH_CTA32=1 timeout 300 python bench.py --no-bkz --no-cpu-baseline > $O/bench_cta32.json 2> $O/bench_cta32.err
B200_UPD_TMA=1 timeout 300 python bench.py --no-bkz --no-cpu-baseline > $O/bench_tma.json 2> $O/bench_tma.err
python - <<'EOF'
import json
for f in ("default", "cta32", "tma"):
    try:
        j = json.loads(open("gpurun_out/r2/bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(j["value"]), "frac", round(j["roofline"]["frac"], 3), "e2e", round(j["e2e"]["value"]),
              "hh", j.get("householder", {}).get("frac_of_hbm_peak"), "enum", j.get("enum", {}).get("gpu_nodes_per_s"))
    except Exception as ex:
        print(f, "ERR", ex)
EOF
echo done
