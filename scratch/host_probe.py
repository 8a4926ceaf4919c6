import os, sys, time, cProfile, pstats
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
X0 = Xh.to(dev); X = torch.empty_like(X0)
def step():
    X.copy_(X0)
    return engine.analysis_device(X, nlev, grid, obs, engine.LOC_GC)
for i in range(3): step()
torch.cuda.synchronize()
for i in range(4):
    t0=time.perf_counter(); r=step(); torch.cuda.synchronize(); t1=time.perf_counter()
    print('wall', round((t1-t0)*1e3,1), 'phases', round(sum(r.ms.values()),1))
pr = cProfile.Profile(); pr.enable()
for i in range(3): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
