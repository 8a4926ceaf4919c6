#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for cfg in "EXB_S2_L2=0 EXB_S2_CAP=2048" "EXB_S2_L2=2 EXB_S2_CAP=2048" "EXB_S2_L2=2 EXB_S2_CAP=1024"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-api 2>&1 | grep -E "^\{|s2\] L2" | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items() if k in ('obs_solve','state_update')})
    else: print(l.strip())" | tail -2
  env $cfg timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:state_sweep_2p -s 2 -c 1 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
