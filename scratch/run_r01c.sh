#!/bin/bash
# round-1 checkpoint run: GPU tests, default bench, launch list, ncu captures of the two dominant kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r01c_pytest.log
python bench.py > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r01c_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv $B > gpurun_out/r01c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:state_update_mma -s 3 -c 1 -f -o gpurun_out/r01c_su_mma $B > gpurun_out/r01c_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:obs_solve_persistent -s 3 -c 1 -f -o gpurun_out/r01c_obs $B > gpurun_out/r01c_ncu3.log 2>&1
tail -3 gpurun_out/r01c_pytest.log; cat gpurun_out/r01c_bench.json | cut -c1-1500
