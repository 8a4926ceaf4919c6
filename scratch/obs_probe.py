# timing of the obs-space solve variants on config-3-like obs (no big state needed: Yp synthesised)
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efa_xray_b200 import engine, _lib
from efa_xray_b200.synth import draw_obs_locations

def run(nobs, nens, cutoff, impls, dtype='f64', reps=3):
    rng = np.random.default_rng(0)
    lat, lon = draw_obs_locations(rng, nobs, 721, 1440)
    dev = torch.device('cuda', 0)
    obs = engine.ObsArrays(value=rng.normal(0, 1, nobs), error=np.ones(nobs), lat=lat, lon=lon,
                           halfwidth=np.full(nobs, cutoff / 2), assimilate=np.ones(nobs, np.uint8),
                           row0=np.zeros(nobs, np.int64), row1=np.zeros(nobs, np.int64), tw0=np.ones(nobs), tw1=np.zeros(nobs))
    # smooth random field sampled at the obs: low-wavenumber waves + noise, like synth.make_case
    lam, phi = np.radians(lon), np.radians(lat)
    amp = rng.normal(0, 1, (6, nens))
    ks = [(1, 1), (2, 1), (3, 2), (4, 3), (2, 3), (5, 2)]
    B = np.stack([np.cos(k * lam + 0.3 * i) * np.cos(l * phi + 0.1 * i) * np.cos(phi) for i, (k, l) in enumerate(ks)], 1)
    Y = B @ amp + 0.3 * rng.standard_normal((nobs, nens))
    tdt = torch.float64 if dtype == 'f64' else torch.float32
    Yp0 = torch.as_tensor(Y - Y.mean(1, keepdims=True)).to(dev).to(tdt)
    Ym0 = torch.as_tensor(Y.mean(1)).to(dev).to(tdt)
    obs_dev, geo = engine.upload_obs(obs, dev, 1)
    out = {}
    res = {}
    for impl in impls:
        os.environ['EXB_OBS_IMPL'] = impl
        ts = []
        for r in range(reps):
            ym, yp = Ym0.clone(), Yp0.clone()
            rec = torch.empty((8, nobs), dtype=torch.float64, device=dev)
            cnt = torch.zeros(2, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.obs_solve(ym, yp, obs_dev, geo, nens, 1, rec, cnt, dtype)
            e1.record(); torch.cuda.synchronize()
            _lib.call('exb_obs_solve_async_status')
            ts.append(e0.elapsed_time(e1))
        out[impl] = dict(ms=ts, pairs=int(cnt[0].item()))
        res[impl] = (ym.cpu().numpy(), yp.cpu().numpy(), rec.cpu().numpy())
    if len(impls) > 1:
        a, b = res[impls[0]], res[impls[1]]
        out['maxdiff_yp'] = float(np.abs(a[1] - b[1]).max() / np.abs(a[1]).max())
        out['maxdiff_rec'] = float(np.nanmax(np.abs(a[2][:7] - b[2][:7]) / (np.abs(a[2][:7]) + 1e-30)))
    return out

if __name__ == '__main__':
    nobs = int(sys.argv[1]); nens = int(sys.argv[2]); cutoff = float(sys.argv[3]); impls = sys.argv[4].split(',')
    dtype = sys.argv[5] if len(sys.argv) > 5 else 'f64'
    print(json.dumps(dict(nobs=nobs, nens=nens, cutoff=cutoff, dtype=dtype, **run(nobs, nens, cutoff, impls, dtype))))
