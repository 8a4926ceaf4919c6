#!/bin/bash
# round 2 checkpoint: full GPU tests, bench (with e2e, api, cpu baseline), reference arm, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
echo "== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu --durations=8 > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 $O/r02n_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02n_bench.json 2> $O/r02n_bench.err; echo "rc=$?"; tail -2 $O/r02n_bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02n_ref.json 2> $O/r02n_ref.err; echo "rc=$?"
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-api > $O/r02n_ncu1.log 2>&1; echo "rc=$?"
echo "== ncu full sweep"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:state_sweep_2p -s 1 -c 1 -o $O/r02n_sweep2p -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api > $O/r02n_ncu2.log 2>&1; echo "rc=$?"
echo "== ncu full dag"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:dag_solve -s 1 -c 1 -o $O/r02n_dag -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api > $O/r02n_ncu3.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench.json') if l.startswith('{')][-1])
print('ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64', round(d['roofline_fp64']['frac'],3))
print('e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'api', d['e2e_api'] and round(d['e2e_api']['ms_per_call'],1), d['e2e_api'] and d['e2e_api']['wall_ms_all_calls'], 'cpu', d.get('cpu_baseline'))
PY
