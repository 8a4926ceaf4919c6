#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for cfg in "EXB_S2_L2=1 EXB_S2_CAP=2048" "EXB_S2_L2=1 EXB_S2_CAP=1024"; do
  echo "== $cfg"
  env $cfg timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:state_sweep_2p -s 1 -c 1 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
