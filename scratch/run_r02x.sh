#!/bin/bash
# C host entry with plans: parity + timing of the host-buffer call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_buffer or c_abi or e2e or golden" > gpurun_out/r02x_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02x_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02x_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e'], d.get('phases_ms'))
PY
