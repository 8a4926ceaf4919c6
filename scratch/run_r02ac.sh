#!/bin/bash
# last check of the final tree: parity tests of the default path and its variants, smoke, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden or host_buffer or edge or one_dim or empty or demo or sharded" > $O/r02ac_quick.log 2>&1; echo "quick rc=$?"; tail -3 $O/r02ac_quick.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline > $O/r02ac_bench.json 2> $O/r02ac_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ac_bench.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64', round(d['roofline_fp64']['frac'],3))
print('e2e', round(d['e2e']['ms_per_step'],1), 'api', d['e2e_api'] and round(d['e2e_api']['ms_per_call'],1), 'launches', d['gpu_launches'])
PY
