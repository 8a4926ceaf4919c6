#!/bin/bash
# order of the two plans: solve's lists filled before the sweep plan is started (default) against the earlier order
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for o in 1 0 1 0; do
EXB_PLAN_ORDER=$o timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api > $O/r02aa_$o.json 2> $O/r02aa_$o.err; echo "order=$o rc=$?"
python - $o <<'PY'
import json,sys
for l in open('gpurun_out/r02aa_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()})
PY
done
