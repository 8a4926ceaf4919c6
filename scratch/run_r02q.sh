#!/bin/bash
# hop latency of the obs-space solve: tiny ensembles make the kernel chain-bound (same geometry, same 3864-hop chain)
cd "$GRAFT_REPO_ROOT" || exit 1
for nens in 8 32 100; do python scratch/obs_probe.py 100000 $nens 2000 dag; done
python scratch/obs_probe.py 100000 8 500 dag
python scratch/obs_probe.py 100000 8 5000 dag
