#!/bin/bash
for d in 0 1; do
EXB_SP_SPLIT=$d EXB_NO_CLOCKS=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('split $d', d['phases_ms']['state_update'])"
done
EXB_SP_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants" 2>&1 | tail -2
