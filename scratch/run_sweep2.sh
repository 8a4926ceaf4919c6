#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or oracle" 2>&1 | tail -2
B="python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e"
EXB_NO_CLOCKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:state_sweep_pipe -s 2 -c 1 -f -o gpurun_out/r01f_sweep_pipe $B > gpurun_out/r01f_ncu.log 2>&1
tail -1 gpurun_out/r01f_ncu.log
