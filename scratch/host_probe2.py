import os, sys, time
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
log = []
orig = _lib.call
def traced(name, *args):
    t0 = time.perf_counter(); r = orig(name, *args); log.append((name, (time.perf_counter()-t0)*1e3)); return r
_lib.call = traced; engine._lib.call = traced
import gc
for i in range(12):
    log.clear()
    t0=time.perf_counter()
    res = engine.analysis_host(Xh, nlev, None, None, obs, engine.LOC_GC, device=dev, dtype=torch.float64, grid=grid, out=Oh)
    torch.cuda.synchronize(); w=(time.perf_counter()-t0)*1e3
    top = sorted(log, key=lambda x:-x[1])[:4]
    print(i, 'wall', round(w,1), 'sum lib calls', round(sum(x[1] for x in log),1), [(n, round(t,1)) for n,t in top], 'gc', gc.get_count())
