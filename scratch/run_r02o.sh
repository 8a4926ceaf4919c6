#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
for cap in 2048 1536 1024; do
  echo "== cap $cap"
  EXB_S2_CAP=$cap timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-api 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items() if k in ('obs_solve','state_update')})"
  EXB_S2_CAP=$cap timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:state_sweep_2p\|dag_solve -s 2 -c 2 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-160
done
