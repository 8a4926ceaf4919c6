#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "obs_solve" 2>&1 | tail -3
timeout 200 python scratch/obs_probe.py 100000 100 2000 dag,persistent 2>&1 | tail -1
timeout 200 python scratch/obs_probe.py 100000 100 500 dag 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dag_ -c 5 --csv python scratch/obs_probe.py 100000 100 2000 dag 2>&1 | grep -v "^==" | awk -F\",\" "{print \$5, \$NF}" | tail -5
