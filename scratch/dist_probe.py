# torchrun --nproc-per-node N scratch/dist_probe.py NOBS NENS CUTOFF : distributed vs replicated obs-space solve
import os, sys, time, json
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efa_xray_b200 import engine, _lib
from efa_xray_b200.synth import draw_obs_locations

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
nobs, nens, cutoff = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
rng = np.random.default_rng(0)
lat, lon = draw_obs_locations(rng, nobs, 721, 1440)
assim = (rng.uniform(0, 1, nobs) > 0.03).astype(np.uint8)
obs = engine.ObsArrays(value=rng.normal(0, 1, nobs), error=np.ones(nobs), lat=lat, lon=lon, halfwidth=np.full(nobs, cutoff / 2),
                       assimilate=assim, row0=np.zeros(nobs, np.int64), row1=np.zeros(nobs, np.int64), tw0=np.ones(nobs), tw1=np.zeros(nobs))
lam, phi = np.radians(lon), np.radians(lat)
amp = rng.normal(0, 1, (6, nens))
ks = [(1, 1), (2, 1), (3, 2), (4, 3), (2, 3), (5, 2)]
B = np.stack([np.cos(k * lam + 0.3 * i) * np.cos(l * phi + 0.1 * i) * np.cos(phi) for i, (k, l) in enumerate(ks)], 1)
Y = B @ amp + 0.3 * rng.standard_normal((nobs, nens))
Yp0 = torch.as_tensor(Y - Y.mean(1, keepdims=True)).to(dev); Ym0 = torch.as_tensor(Y.mean(1)).to(dev)
obs_dev, geo = engine.upload_obs(obs, dev, 1)

def run(distributed):
    ym, yp = Ym0.clone(), Yp0.clone()
    rec = torch.empty((8, nobs), dtype=torch.float64, device=dev); cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    plan = engine.ObsPlan(obs_dev, geo, nobs, 1, rank, world) if distributed else engine.ObsPlan(obs_dev, geo, nobs, 1)
    plan.finish()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if distributed:
        ok = engine.obs_solve_distributed(ym, yp, obs_dev, geo, nens, 1, rec, cnt, 'f64', plan)
        assert ok
    else:
        engine.obs_solve(ym, yp, obs_dev, geo, nens, 1, rec, cnt, 'f64', plan=plan)
    e1.record(); torch.cuda.synchronize()
    _lib.call('exb_obs_solve_async_status')
    plan.destroy()
    return e0.elapsed_time(e1), ym.cpu().numpy(), yp.cpu().numpy(), rec.cpu().numpy(), int(cnt[0].item())

ref = run(False)
out = {}
for rep in range(3):
    a = run(True); b = run(False)
    out.setdefault('dist_ms', []).append(round(a[0], 2)); out.setdefault('repl_ms', []).append(round(b[0], 2))
scale = np.abs(ref[2]).max()
out['pairs'] = (a[4], ref[4])
out['maxdiff_yp'] = float(np.abs(a[2] - ref[2]).max() / scale)
out['maxdiff_ym'] = float(np.abs(a[1] - ref[1]).max())
m = ~np.isnan(ref[3])
out['nan_pattern_equal'] = bool((np.isnan(a[3]) == np.isnan(ref[3])).all())
out['maxdiff_rec'] = float(np.max(np.abs(a[3][m] - ref[3][m]) / (np.abs(ref[3][m]) + 1e-30)))
if rank == 0:
    print(json.dumps(dict(nobs=nobs, nens=nens, cutoff=cutoff, world=world, **out)))
dist.destroy_process_group()
