#!/bin/bash
O=gpurun_out/s2
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_update_row_stream --launch-skip 3 --launch-count 1 -f -o $O/stream_v1 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 > $O/ncu_stream_v1.log 2>&1
tail -3 $O/ncu_stream_v1.log
ls -la $O
