#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > $O/r02p_n2.json 2> $O/r02p_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02p_n2.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], 'steps', [round(x,1) for x in d['step_ms']], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['sharded_check'] and d['sharded_check']['ok'])
PY
