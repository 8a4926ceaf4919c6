import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from efa_xray_b200 import engine, synth, _lib
class A: pass
a = A(); a.config='config3'; a.nobs=int(os.environ.get('NOBS','100000')); a.cutoff_km=2000.0; a.seed=0
cfg = dict(synth.CONFIGS['config3']); cfg['nobs']=a.nobs
nlev=3; ny,nx,nens=cfg['ny'],cfg['nx'],cfg['nmem']
Xh = torch.empty((nlev*ny*nx, nens), dtype=torch.float64).pin_memory()
case,_ = bench.build_case(a, out=Xh.numpy().reshape(3,1,ny,nx,nens))
obs = bench.obs_arrays(case)
dev = torch.device('cuda',0)
grid = engine.GridTables(case.lat2d, case.lon2d, dev)
X0 = Xh.to(dev); X = torch.empty_like(X0)
for it in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    X.copy_(X0); torch.cuda.synchronize(); t1=time.perf_counter()
    res = engine.analysis_device(X, nlev, grid, obs, 1)
    torch.cuda.synchronize(); t2=time.perf_counter()
    print('iter',it,'copy %.1f ms  analysis wall %.1f ms  phases'%(1e3*(t1-t0),1e3*(t2-t1)), {k:round(v,1) for k,v in res.ms.items()}, flush=True)
# finer: time obs_solve call on host
torch.cuda.synchronize()
import ctypes as C
sfx='f64'
obs_dev, geo = engine.upload_obs(obs, dev, 1)
Yp, nex = engine.ob_priors(X0, grid, obs, sfx)
Ym = torch.empty(obs.nobs, dtype=torch.float64, device=dev)
_lib.call('exb_split_mean_pert_f64', _lib.ptr(Yp), _lib.ptr(Ym), obs.nobs, nens, _lib.stream_ptr())
rec = torch.empty((8, obs.nobs), dtype=torch.float64, device=dev); cnt=torch.zeros(2,dtype=torch.int64,device=dev)
for it in range(3):
    Y2=Yp.clone(); M2=Ym.clone(); torch.cuda.synchronize(); t0=time.perf_counter()
    engine.obs_solve(M2, Y2, obs_dev, geo, nens, 1, rec, cnt, sfx); t1=time.perf_counter()
    torch.cuda.synchronize(); t2=time.perf_counter()
    print('obs_solve host call %.1f ms, total %.1f ms'%(1e3*(t1-t0),1e3*(t2-t0)), flush=True)
