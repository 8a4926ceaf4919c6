#!/bin/bash
# BASELINE configs 1, 2, 4 and the config-5 sweep (observation count x localisation cutoff) at N = 1.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
OUT=$O/r02_configs.jsonl; : > $OUT
run() { echo "== $*" >&2; timeout 900 python bench.py --steps 3 --warmup 3 --no-api "$@" >> $OUT 2>> $O/r02_configs.err || echo "{\"failed\": \"$*\"}" >> $OUT; }
run --config config1 --cutoff-km 2000 --cpu-seconds 25
run --config config2 --cutoff-km 2000 --cpu-seconds 15
run --config config4 --cutoff-km 2000 --no-cpu-baseline
run --config config4 --cutoff-km 2000 --dtype f32 --no-cpu-baseline
for nobs in 1000 10000 100000 1000000; do
  for cut in 500 1000 2000 5000; do
    if [ "$nobs" = "100000" ] && [ "$cut" = "2000" ]; then continue; fi     # the headline line (BENCH)
    run --config config3 --nobs $nobs --cutoff-km $cut --no-cpu-baseline --no-e2e
  done
done
# config 1 in full on the host: the oracle's complete serial loop (the reference's CPU cost for its own demo-sized case)
python - >> $OUT 2>> $O/r02_configs.err <<'PY'
import json, os, time, sys
sys.path.insert(0, os.getcwd())
from efa_xray_b200 import synth
from oracle import ensrf_oracle as O
case = synth.make_case(cutoff_km=2000.0, seed=0, **synth.CONFIGS['config1'])
st, obs = O.State.from_case(case), O.obs_from_case(case)
t0 = time.perf_counter()
O.ensrf_update(st, obs, loc='GC')
dt = time.perf_counter() - t0
print(json.dumps({'cpu_full_run': 'config1', 'seconds': dt, 'obs': len(obs), 'ms_per_ob': 1e3 * dt / len(obs), 'cores': os.cpu_count()}))
PY
wc -l $OUT
