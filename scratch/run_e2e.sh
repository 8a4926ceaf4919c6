#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
EXB_NO_CLOCKS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_bench.json 2> gpurun_out/e2e_bench.err
tail -3 gpurun_out/e2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/e2e_bench.json'))
print(d['ms_per_step'], d['step_ms'], d['phases_ms'], d['e2e'], d['gpu_launches'])
PY
