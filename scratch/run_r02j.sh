#!/bin/bash
# N = 2: distributed obs-space solve with round-robin and block dealing, sharded check, multi-GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -3
echo "== multi-gpu tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/r02j_multi.log 2>&1; echo "rc=$?"; tail -3 $O/r02j_multi.log
for blk in 1 64 256 1024; do
  echo "== N=2 block $blk"
  EXB_OBS_DIST_BLOCK=$blk timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02j_n2_b$blk.json 2> $O/r02j_n2_b$blk.err; echo "rc=$?"
done
EXB_OBS_DIST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02j_n2_repl.json 2> $O/r02j_n2_repl.err; echo "repl rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02j_n2_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, d['config']['obs_solve'], d.get('sharded_check',{}) and d['sharded_check'].get('ok'), d['sharded_check'] and '%.1e'%d['sharded_check']['rel_to_increment'])
    except Exception as e:
        print(f, 'failed', e)
PY
