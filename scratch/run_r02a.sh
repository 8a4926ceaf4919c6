#!/bin/bash
# round 2, first GPU checkpoint: new two-phase sweep kernel vs the round-1 kernel, full GPU test suite, bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r02a_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 $O/r02a_smoke.log
echo "== sweep variants (quick)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_variants or golden" > $O/r02a_quick.log 2>&1; echo "quick rc=$?"; tail -5 $O/r02a_quick.log
echo "== bench v3"; EXB_SP_IMPL=v3 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02a_bench_v3.json 2> $O/r02a_bench_v3.err; echo "rc=$?"
echo "== bench 2p"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02a_bench_2p.json 2> $O/r02a_bench_2p.err; echo "rc=$?"
python - <<'PY'
import json
for n in ('v3','2p'):
    try:
        d=json.load(open('gpurun_out/r02a_bench_%s.json'%n))
        print(n, 'ms', round(d['ms_per_step'],2), 'phases', {k:round(v,2) for k,v in d['phases_ms'].items()}, 'fp64 frac', round(d['roofline_fp64']['frac'],3))
    except Exception as e:
        print(n, 'failed', e)
PY
echo "== full gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02a_pytest.log
echo "== bench full"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "rc=$?"; tail -3 $O/r02a_bench.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02a_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02a_ncu1.log 2>&1; echo "rc=$?"
echo "== ncu full 2p"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:state_sweep_2p -s 2 -c 1 -o $O/r02a_sweep2p -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02a_ncu2.log 2>&1; echo "rc=$?"
ls -la $O | tail -15
