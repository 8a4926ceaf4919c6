#!/bin/bash
for d in 0 16 32 17 80 81 1 64; do
EXB_SP_DBG=$d EXB_NO_CLOCKS=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dbg $d', d['phases_ms']['state_update'])"
done
