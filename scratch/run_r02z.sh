#!/bin/bash
# final checkpoint of round 2 (after the C-entry plans and the kernel templating): what the driver runs at round end
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== full gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02z_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py > $O/r02z_bench.json 2> $O/r02z_bench.err; echo "rc=$?"; tail -2 $O/r02z_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference > $O/r02z_ref.json 2> $O/r02z_ref.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02z_bench.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64', round(d['roofline_fp64']['frac'],3), 'hbm frac', round(d['roofline']['frac'],2), 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'api', d['e2e_api'] and round(d['e2e_api']['ms_per_call'],1), 'cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
r=json.loads([l for l in open('gpurun_out/r02z_ref.json') if l.startswith('{')][-1])
print('ref', r['value'], r['ms_per_step'])
PY
