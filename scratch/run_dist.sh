#!/bin/bash
export EXB_NO_CLOCKS=1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "obs_solve or oracle or config3" 2>&1 | tail -2
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scratch/dist_probe.py 3000 50 2500 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scratch/dist_probe.py 100000 100 2000 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM\|NCCL version\|^$" | tail -4
scratch/run_n.sh 2
