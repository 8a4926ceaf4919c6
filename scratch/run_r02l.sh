#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== ubench"; (cd scratch/ubench && ./consumer_ubench) | tee $O/r02l_consumer_ubench.txt
echo "== quick tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden or obs_solve or edge or config1 or dag_solve" > $O/r02l_quick.log 2>&1; echo "quick rc=$?"; tail -4 $O/r02l_quick.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-api > $O/r02l_bench.json 2> $O/r02l_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02l_bench.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), 'phases', {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64 frac', round(d['roofline_fp64']['frac'],3))
PY
for dbg in 0 1 2 4 6; do echo "== prof dbg=$dbg"; EXB_S2_DEBUG=$dbg EXB_S2_PROF=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api 2>&1 | grep "s2 prof" | tail -1; done
