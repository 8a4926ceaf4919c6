#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or oracle or config3 or c_abi" 2>&1 | tail -3
EXB_NO_CLOCKS=1 timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep1_bench.json 2> gpurun_out/sweep1_bench.err
tail -3 gpurun_out/sweep1_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep1_bench.json'))
print(d['ms_per_step'], d['phases_ms'], d['roofline_fp64']['achieved'], d['e2e']['ms_per_step'])
PY
