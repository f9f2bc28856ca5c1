#!/bin/bash
echo "== BKZ + enum tests"
timeout 900 python -m pytest tests/test_bkz_gpu.py tests/test_enum_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
echo "== BKZ-60 tours, fixed region (default): seeds 2 2 1 3"
timeout 400 python tools/gpurun_bkz_seed.py 1000 2 2 1 3 2>&1 | grep seed
echo "== shrinking radius: seeds 2 2"
B200_BKZ_SHRINK=1 timeout 200 python tools/gpurun_bkz_seed.py 1000 2 2 2>&1 | grep seed
echo done
