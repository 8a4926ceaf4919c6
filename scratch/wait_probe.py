import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efa_xray_b200 import engine, synth, _lib
import bench
class A: pass
a = A(); a.config='config3'; a.nobs=None; a.cutoff_km=2000.0; a.seed=0
cfg = dict(synth.CONFIGS['config3']); nlev=3; ny,nx,nens=cfg['ny'],cfg['nx'],cfg['nmem']
Xh = torch.empty((nlev*ny*nx, nens), dtype=torch.float64)
case,_ = bench.build_case(a, out=Xh.numpy().reshape(3,1,ny,nx,nens))
obs = bench.obs_arrays(case)
dev = torch.device('cuda',0)
grid = engine.GridTables(case.lat2d, case.lon2d, dev)
X = Xh.to(dev)
# replicate analysis_device but with an 8-slot counter buffer
obs_dev, geo = engine.upload_obs(obs, dev, 1)
Yp, nex = engine.ob_priors(X, grid, obs, 'f64', nlev=nlev)
Ym = torch.empty(obs.nobs, dtype=torch.float64, device=dev)
_lib.call('exb_split_mean_pert_f64', _lib.ptr(Yp), _lib.ptr(Ym), obs.nobs, nens, _lib.stream_ptr())
rec = torch.empty((8, obs.nobs), dtype=torch.float64, device=dev)
cnt = torch.zeros(8, dtype=torch.int64, device=dev)
engine.obs_solve(Ym, Yp, obs_dev, geo, nens, 1, rec, cnt, 'f64')
for rep in range(2):
    cnt.zero_(); Xc = X.clone(); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.state_sweep_fused(Xc, nlev, ny, nx, grid.u, Yp, rec, geo, obs.nobs, 1, cnt); e1.record(); torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    print('ms', e0.elapsed_time(e1))
