#!/bin/bash
# streaming update kernel: parity, then timing against the register-staged kernel
O=gpurun_out/s2
mkdir -p $O
echo "== parity"
timeout 600 python -m pytest tests/test_gso_gpu.py -m gpu -x -q -p no:cacheprovider -k "streaming" 2>&1 | tail -15
echo "== timing"
for W in ${WLIST:-10 8 5}; do
echo "-- stream NW=$W"
B200_ST_WARPS=$W timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 100 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); print('value', round(j['value'],1), 'kernel', round(j['roofline']['achieved'],1), 'frac', round(j['roofline']['frac'],4), 'ms', round(j['roofline']['ms_per_launch'],4), 'e2e', round(j['e2e']['value'],1))
    elif 'rror' in ln: print(ln.strip()[:300])
"
done
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_update_row_stream --launch-skip 3 --launch-count 1 -f -o $O/stream_$NCU python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_stream_$NCU.log 2>&1
tail -2 $O/ncu_stream_$NCU.log | cut -c1-200
fi
echo done
