#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== quick tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden or fp32 or edge or demo or one_dim" > $O/r02d_quick.log 2>&1; echo "quick rc=$?"; tail -8 $O/r02d_quick.log
echo "== bench 2p (prof)"; EXB_S2_PROF=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-api > $O/r02d_bench_2p.json 2> $O/r02d_bench_2p.err; echo "rc=$?"; tail -2 $O/r02d_bench_2p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench_2p.json'))
print('2p ms', round(d['ms_per_step'],2), 'phases', {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64 frac', round(d['roofline_fp64']['frac'],3))
PY
