#!/bin/bash
# two GPUs: sharded single-block enumeration over CUDA IPC (shared ticket) through bench.py under torchrun, no BKZ
O=gpurun_out/s2
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu-baseline --no-bkz > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err | cut -c1-300
python -c "
import json; j=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]); print('value', j['value'], 'enum', {k:v for k,v in j.get('enum',{}).items() if k!='workload'})"
