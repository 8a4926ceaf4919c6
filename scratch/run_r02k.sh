#!/bin/bash
# N = 8: distributed obs-space solve: round-robin vs block dealing vs replicated
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
echo "== multi-gpu tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/r02k_multi.log 2>&1; echo "rc=$?"; tail -3 $O/r02k_multi.log
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02k_n8_$tag.json 2> $O/r02k_n8_$tag.err; echo "$tag rc=$?"; }
run b256 EXB_OBS_DIST_BLOCK=256
run b1 EXB_OBS_DIST_BLOCK=1
run b1024 EXB_OBS_DIST_BLOCK=1024
run repl EXB_OBS_DIST=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02k_n8_*.json')):
    try:
        line=[l for l in open(f) if l.startswith('{')][-1]
        d=json.loads(line)
        sc=d.get('sharded_check') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], sc.get('ok'), '%.1e'%sc.get('rel_to_increment',-1))
    except Exception as e:
        print(f, 'failed', e)
PY
