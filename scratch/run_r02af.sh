#!/bin/bash
# full-size parity tests that go through the sweep plan, on the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 85 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config3_full or config1_full or config4_full or config2_full or logical_ranks_config3" > $O/r02af_full.log 2>&1; echo "rc=$?"; tail -3 $O/r02af_full.log
