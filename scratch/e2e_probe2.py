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
X0 = Xh.to(dev)
main = torch.cuda.current_stream()
copy_out = torch.cuda.Stream(device=dev)
npts = ny*nx
O3 = Oh.view(nlev, npts, nens)
for mode in ('nodl', 'dl', 'dl_after_all', 'nodl'):
    X = X0.clone(); X3 = X.view(nlev, npts, nens)
    marks = []
    def cb(ya, yb):
        e = torch.cuda.Event(enable_timing=True); e.record(main); marks.append(e)
        if mode == 'dl':
            copy_out.wait_event(e)
            with torch.cuda.stream(copy_out):
                for lev in range(nlev):
                    O3[lev, ya*nx:yb*nx].copy_(X3[lev, ya*nx:yb*nx], non_blocking=True)
    bands = engine.sweep_band_schedule(ny)
    torch.cuda.synchronize()
    res = engine.analysis_device(X, nlev, grid, obs, engine.LOC_GC, sweep_bands=bands, on_band_done=cb)
    torch.cuda.synchronize()
    per = [round(marks[i].elapsed_time(marks[i+1]),1) for i in range(len(marks)-1)]
    print(mode, 'state_update', round(res.ms['state_update'],1), 'per-band (from 2nd)', per)
