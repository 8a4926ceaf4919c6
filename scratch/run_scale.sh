#!/bin/bash
# strong scaling of one analysis (config 3), launched exactly as the driver does; usage: run_scale.sh "2 4" | "8" [sweep]
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for n in $1; do
  echo "== N=$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 5 --warmup 3 > $O/r02_scale_n$n.json 2> $O/r02_scale_n$n.err; echo "rc=$?"
done
if [ "$2" = "sweep" ]; then
  n=8
  for cut in 500 1000 5000; do
    echo "== N=8 cutoff $cut"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 3 --warmup 3 --cutoff-km $cut --no-e2e --no-cpu-baseline > $O/r02_scale_n${n}_c$cut.json 2> $O/r02_scale_n${n}_c$cut.err; echo "rc=$?"
  done
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_scale_n*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        sc=d.get('sharded_check') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'check', sc.get('ok'), d['config']['bands'])
    except Exception as e:
        print(f, 'failed', e)
PY
