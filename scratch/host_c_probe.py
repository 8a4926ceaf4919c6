# exb_ensrf_host_f64 on config 3 with a pinned buffer: wall time and stats
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efa_xray_b200 import engine, synth, _lib
import bench
class A: pass
a = A(); a.config='config3'; a.nobs=None; a.cutoff_km=2000.0; a.seed=0
cfg = dict(synth.CONFIGS['config3']); nlev=3; ny,nx,nens=cfg['ny'],cfg['nx'],cfg['nmem']
Xh = torch.empty((nlev*ny*nx, nens), dtype=torch.float64).pin_memory()
case,_ = bench.build_case(a, out=Xh.numpy().reshape(3,1,ny,nx,nens))
X0 = Xh.clone()
obs = bench.obs_arrays(case)
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
lat, lon = np.ascontiguousarray(case.lat2d), np.ascontiguousarray(case.lon2d)
diag = np.zeros((4, obs.nobs)); stats = np.zeros(8)
arrs = [np.ascontiguousarray(x) for x in (obs.value, obs.error, obs.lat, obs.lon, obs.halfwidth, obs.assimilate, obs.row0, obs.row1, obs.tw0, obs.tw1)]
for rep in range(4):
    Xh.copy_(X0)
    t0 = time.perf_counter()
    _lib.call('exb_ensrf_host_f64', C.c_void_p(Xh.data_ptr()), nlev, ny, nx, nens, ptr(lat), ptr(lon), obs.nobs, *[ptr(x) for x in arrs], 1, 1.0, ptr(diag), ptr(stats))
    print('wall ms', round((time.perf_counter()-t0)*1e3, 1), 'stats', np.round(stats, 1))
# agreement with the Python pipeline
Oh = torch.empty_like(Xh).pin_memory()
grid = engine.GridTables(case.lat2d, case.lon2d, torch.device('cuda', 0))
res = engine.analysis_host(X0, nlev, None, None, obs, 1, device='cuda:0', dtype=torch.float64, grid=grid, out=Oh)
print('max abs diff C entry vs python pipeline', float((Oh - Xh).abs().max()), 'diag diff', float(np.nanmax(np.abs(diag[3] - res.post_var))))
