#!/bin/bash
export EXB_NO_CLOCKS=1
for cfg in "1 3000" "7 3000" "64 20000"; do set -- $cfg
EXB_OBS_DIST_BLOCK=$1 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scratch/dist_probe.py $2 50 2500 > gpurun_out/dist_dbg_$1.log 2>&1
echo "blk=$1 nobs=$2:"; grep -h "ExbError\|AssertionError\|^{" gpurun_out/dist_dbg_$1.log | head -2 | cut -c1-330
done
