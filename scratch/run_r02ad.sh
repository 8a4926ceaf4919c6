#!/bin/bash
# launch list of the final build (after the plain run of the same command has exited 0)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-api > $O/r02ad_plain.json 2> $O/r02ad_plain.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02ad_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-api > $O/r02ad_ncu1.log 2>&1; echo "ncu rc=$?"
wc -l $O/r02ad_launches.csv
