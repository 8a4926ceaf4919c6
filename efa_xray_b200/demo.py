"""The single-point EFA demonstration of efa_demo.ipynb (cell 11, `enkf`, and cell 14) on the GPU path.

The notebook updates a forecast trajectory at ONE point -- a [nTimes, nEnsMems] array of one variable -- with
observations of its first few valid times, in random order, without localisation.  Here the same update runs through
the library: the trajectory is a state with nTimes levels on a 1 x 1 grid, every observation's forward operator is
the one-hot row select of the notebook (cell 11, lines 55-59), and the serial loop is the obs-space solve + state
sweep of the EnSRF path with loc = None.
"""
from __future__ import print_function

import numpy as np

from . import engine, _lib


def enkf(obs, prior, obs_range=(1, 2), ob_error=1.0, inflation=1.0, order=None, seed=None, point=(47.4489, -122.3094)):
    """Updates a prior ensemble given a selection of obs (efa_demo.ipynb cell 11).

    obs        list of observation values, one per successive valid time, starting at the first valid time
    prior      [nTimes, nEnsMems] array of the state variable
    obs_range  (start, end), 1-based inclusive: obs[start-1:end] are assimilated
    ob_error   error variance of every observation
    inflation  multiplicative inflation of the prior perturbations (cell 11, lines 40-42)
    order      the order in which the selected obs are assimilated (a permutation of range(n)); the notebook shuffles
               them on every call (lines 44-46): default = a random permutation drawn from `seed`
    Returns the posterior [nTimes, nEnsMems] (mean + perturbations), like the notebook."""
    import torch
    _lib.require_device()
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    nt, nens = prior.shape
    sel = list(obs[obs_range[0] - 1:obs_range[-1]])
    n = len(sel)
    if order is None:
        order = np.random.default_rng(seed).permutation(n)
    order = np.asarray(order, dtype=np.int64)
    assert sorted(order.tolist()) == list(range(n)), 'order must be a permutation of the selected obs'
    ob_index = obs_range[0] + order - 1                          # cell 11, line 52: row of the state each ob observes
    dev = torch.device('cuda', torch.cuda.current_device())
    X = torch.as_tensor(prior).to(dev)
    if n == 0:
        if inflation != 1.0:
            engine.analysis_device(X, nt, engine.GridTables(np.array([[point[0]]]), np.array([[point[1] % 360.0]]), dev),
                                   _point_obs(np.zeros(0), np.zeros(0, dtype=np.int64), ob_error, point), engine.LOC_NONE,
                                   inflation=np.full(nt, float(inflation)))
        return X.cpu().numpy()
    grid = engine.GridTables(np.array([[point[0]]]), np.array([[point[1] % 360.0]]), dev)
    oa = _point_obs(np.array(sel, dtype=np.float64)[order], ob_index, ob_error, point)
    if inflation != 1.0:
        import ctypes as C
        fac = np.full(nt, float(inflation))
        _lib.call('exb_inflate_f64', _lib.ptr(X), nt, nens, fac.ctypes.data_as(C.c_void_p), nt, 1, _lib.stream_ptr())
    # H = one-hot row select (cell 11, lines 55-59): ye = the state row itself
    Y = X[torch.as_tensor(ob_index, device=dev)].contiguous()
    engine.analysis_device(X, nt, grid, oa, engine.LOC_NONE, Y=Y)
    return X.cpu().numpy()


def _point_obs(values, ob_index, ob_error, point):
    n = values.shape[0]
    return engine.ObsArrays(value=values, error=np.full(n, float(ob_error)), lat=np.full(n, float(point[0])),
                            lon=np.full(n, float(point[1] % 360.0)), halfwidth=np.ones(n), assimilate=np.ones(n, dtype=np.uint8),
                            row0=ob_index.astype(np.int64), row1=ob_index.astype(np.int64), tw0=np.ones(n), tw1=np.zeros(n))


def synthetic_point_ensemble(ntimes=13, nmems=21, seed=0):
    """A GEFS-like 2 m temperature trajectory at one point (the notebook downloads one with Siphon; there is no network
    here): spread growing with lead time around a diurnal cycle.  Returns (times [ntimes] datetime64, [ntimes, nmems])."""
    rng = np.random.default_rng(seed)
    lead = np.arange(ntimes) * 6.0
    base = 278.0 + 3.0 * np.sin(2 * np.pi * (lead - 9.0) / 24.0) - 0.02 * lead
    drift = np.cumsum(rng.normal(0.0, 0.45, (ntimes, nmems)), axis=0)
    times = np.datetime64('2016-01-15T00:00:00') + (lead * 3600).astype('timedelta64[s]')
    return times, base[:, None] + 0.3 * rng.standard_normal(nmems)[None, :] + drift
