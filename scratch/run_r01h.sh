#!/bin/bash
# round-1 checkpoint v3: GPU tests, default bench, reference arm, launch list, ncu captures of the two dominant kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r01h_pytest.log
timeout 400 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01h_ref.json 2> gpurun_out/r01h_ref.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
EXB_NO_CLOCKS=1 $B > gpurun_out/r01h_plain.log 2>&1 || exit 1
EXB_NO_CLOCKS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01h_launches.csv $B > gpurun_out/r01h_ncu1.log 2>&1
EXB_NO_CLOCKS=1 ncu --set full --clock-control none --import-source on -k regex:state_sweep_pipe -s 3 -c 1 -f -o gpurun_out/r01h_sweep $B > gpurun_out/r01h_ncu2.log 2>&1
EXB_NO_CLOCKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dag_solve -s 3 -c 1 -f -o gpurun_out/r01h_dag $B > gpurun_out/r01h_ncu3.log 2>&1
tail -3 gpurun_out/r01h_pytest.log; cut -c1-1800 gpurun_out/r01h_bench.json; echo; cut -c1-600 gpurun_out/r01h_ref.json
