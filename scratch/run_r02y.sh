#!/bin/bash
# two CTAs of 8 warps per SM (EXB_S2_WARPS=8) against one of 16: parity of the variant, then time and phase clocks
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== quick tests (8 warps)"; EXB_S2_WARPS=8 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden or fp32 or edge or one_dim or multi_chunk or chunk" > $O/r02y_quick.log 2>&1; echo "quick rc=$?"; tail -4 $O/r02y_quick.log
for w in 8 16; do
echo "== bench warps=$w"; EXB_S2_WARPS=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-api > $O/r02y_bench_$w.json 2> $O/r02y_bench_$w.err; echo "rc=$?"
python - $w <<'PY'
import json,sys
for l in open('gpurun_out/r02y_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'phases', {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64 frac', round(d['roofline_fp64']['frac'],3))
PY
done
echo "== prof 8"; EXB_S2_WARPS=8 EXB_S2_PROF=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-api 2>&1 | grep "s2 prof" | tail -1
