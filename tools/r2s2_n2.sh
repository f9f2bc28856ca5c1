#!/bin/bash
# two GPUs: bench.py under torchrun (sharded single-block enumeration over CUDA IPC + BKZ-60 child on 2 devices)
O=gpurun_out/s2
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus ${NG:-2} --steps 50 --warmup 3 --no-cpu-baseline > $O/bench_n${NG:-2}.json 2> $O/bench_n${NG:-2}.err
tail -3 $O/bench_n${NG:-2}.err | cut -c1-300
python -c "
import json; j=json.loads(open('$O/bench_n${NG:-2}.json').read().strip().splitlines()[-1]); print('value', j['value'], 'enum', {k:v for k,v in j.get('enum',{}).items() if k!='workload'}); print('bkz60', {k:v for k,v in (j.get('bkz60') or {}).items() if k not in ('workload',)})"
