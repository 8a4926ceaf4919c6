#!/bin/bash
# usage: run_n.sh N  -> bench at N GPUs, JSON to gpurun_out/scale_nN.json
N=$1
mkdir -p gpurun_out
EXB_NO_CLOCKS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d['config']['bands'])
PY
grep -v "Warning\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/scale_n$N.err | tail -3 | cut -c1-300
