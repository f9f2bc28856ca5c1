#!/bin/bash
# lane-per-lattice update_R (B200_HH_X32=1): parity under a short timeout (a protocol bug would hang), then timing
O=gpurun_out/s2
mkdir -p $O
echo "== x32 parity"
timeout 240 python -m pytest tests/test_hh_gpu.py -m gpu -x -q -p no:cacheprovider -k "lane_per_lattice" 2>&1 | tail -6
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "parity failed or timed out (rc $rc): no timing"; exit 0; fi
for B in 2960 4736; do
for X in 1 0; do
echo "-- batch $B  B200_HH_X32=$X"
B200_BENCH_HH_BATCH=$B B200_HH_X32=$X timeout 300 python bench.py --no-bkz --no-cpu-baseline --no-traffic --steps 20 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'): print({k: v for k, v in json.loads(ln).get('householder', {}).items() if k != 'workload'})"
done
done
echo done
