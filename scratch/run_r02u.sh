#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for cfg in "EXB_DAG_LOCAL_GPU=1" "EXB_DAG_LOCAL_GPU=0" "EXB_DAG_LOCAL_GPU=1 EXB_DAG_HOT=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], d['sharded_check']['ok'])"
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
