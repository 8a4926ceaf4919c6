#!/bin/bash
export EXB_NO_CLOCKS=1
EXB_OBS_DIST_BLOCK=7 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scratch/dist_probe.py 3001 50 2500 > gpurun_out/dist_dbg.log 2>&1
grep -v "Warning\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/dist_dbg.log | grep -B2 -A12 "Traceback" | head -50
tail -3 gpurun_out/dist_dbg.log | cut -c1-300
