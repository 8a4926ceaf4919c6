#!/bin/bash
# sweep plan created from a helper thread (the main thread goes on to the solve): off / on / on and started before the ob priors
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
run() {
  env $1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api > $O/r02ab_$2.json 2> $O/r02ab_$2.err; echo "$1 rc=$?"
  python - $2 <<'PY'
import json,sys
for l in open('gpurun_out/r02ab_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()})
PY
}
run "EXB_PLAN_THREAD=0" t0
run "EXB_PLAN_THREAD=1" t1
run "EXB_PLAN_THREAD=1 EXB_PLAN_EARLY=1" t1e
