import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efa_xray_b200 import engine, synth, _lib
import bench
class A: pass
a = A(); a.config='config3'; a.nobs=None; a.cutoff_km=2000.0; a.seed=0
cfg = dict(synth.CONFIGS['config3']); nlev=3; ny,nx,nens=cfg['ny'],cfg['nx'],cfg['nmem']
Xh = torch.empty((nlev*ny*nx, nens), dtype=torch.float64).pin_memory()
case,_ = bench.build_case(a, out=Xh.numpy().reshape(3,1,ny,nx,nens))
obs = bench.obs_arrays(case)
dev = torch.device('cuda',0)
grid = engine.GridTables(case.lat2d, case.lon2d, dev)
Oh = torch.empty_like(Xh).pin_memory()
# raw copy bandwidth
Xd = torch.empty_like(Xh, device=dev)
for name, fn in (('h2d', lambda: Xd.copy_(Xh, non_blocking=True)), ('d2h', lambda: Oh.copy_(Xd, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); print(name, ms, 'ms', Xh.numel()*8/ms/1e6, 'GB/s')
del Xd
for pipe in (True, False, True):
    t0=time.perf_counter()
    res = engine.analysis_host(Xh, nlev, case.lat2d, case.lon2d, obs, engine.LOC_GC, device=dev, dtype=torch.float64, grid=grid, out=Oh, pipeline=pipe)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print('pipeline', pipe, 'wall ms', (t1-t0)*1e3, {k: round(v,1) for k,v in res.ms.items()})
# banding without downloads vs with
X0 = Xh.to(dev)
for bands in (None, [(0,721)], engine.sweep_band_schedule(721), [(0,88),(88,176),(176,264),(264,352),(352,440),(440,528),(528,616),(616,672),(672,721)]):
    X = X0.clone(); torch.cuda.synchronize(); t0=time.perf_counter()
    res = engine.analysis_device(X, nlev, grid, obs, engine.LOC_GC, sweep_bands=bands)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print('bands', bands, 'wall', round((t1-t0)*1e3,1), {k: round(v,1) for k,v in res.ms.items()})
