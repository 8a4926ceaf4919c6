#!/bin/bash
# tile-list kernel with the record loads unrolled: parity of every sweep variant, then the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden or sharded" > $O/r02ae_quick.log 2>&1; echo "quick rc=$?"; tail -2 $O/r02ae_quick.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api > $O/r02ae_bench.json 2> $O/r02ae_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ae_bench.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()})
PY
