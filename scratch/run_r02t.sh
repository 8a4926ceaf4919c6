#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for hot in 0 1; do echo "== N=1 hot=$hot"; EXB_DAG_HOT=$hot python scratch/obs_probe.py 100000 100 2000 dag; EXB_DAG_HOT=$hot python scratch/obs_probe.py 100000 100 5000 dag; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "obs_solve or dag_solve or golden" 2>&1 | tail -2
for hot in 0 1; do
  echo "== N=2 hot=$hot"
  EXB_DAG_HOT=$hot timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], d['sharded_check']['ok'])"
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
