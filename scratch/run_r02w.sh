#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for cfg in "EXB_DIST_SEARCH=1" "EXB_DIST_SEARCH=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], 'check', d['sharded_check']['ok'], d['sharded_check']['rel_to_increment'], 'e2e', round(d['e2e']['ms_per_step'],1))"
done
