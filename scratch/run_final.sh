#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
scratch/run_n.sh 4
