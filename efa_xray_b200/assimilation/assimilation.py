"""Assimilation base class and the module-level update() entry point
(efa_xray/assimilation/assimilation.py:10-230)."""
from __future__ import print_function

from copy import deepcopy

import numpy as np

import efa_xray_b200 as _pkg
from .. import _lib
from .. import engine


class ObTimeOutsideState(AttributeError, ValueError):
    """An observation's time is outside the state's valid times.  The reference's interpolate returns None
    there (state/ensemble.py:206-208) and compute_ob_priors then dies with AttributeError on `.mean()`
    (assimilation/assimilation.py:47); this is that failure with a message."""


class Assimilation():
    """Computes obs priors and formats the state for the update (assimilation.py:10-171)."""

    def __init__(self, state, obs, nproc=1, inflation=None, verbose=False):
        """inflation: None, a float (all variables), or a dict of {variable name: float}
        (assimilation.py:19-26; the file-name and per-dimension forms need xarray/netCDF and are not ported).
        nproc is accepted and ignored, as in the reference (assimilation.py:31)."""
        self.prior = state
        self.obs = obs
        self.verbose = verbose
        self.nproc = nproc
        self.inflation = inflation
        self.is_inflated = False
        # the reference also deep-copies the state into self.post and never uses it (assimilation.py:29)

    # ---- marshalling ------------------------------------------------------------------
    def _loc_mode(self, loc):
        if loc in [None, False]:
            return engine.LOC_NONE
        if loc == 'GC':
            return engine.LOC_GC
        raise ValueError("loc must be None, False or 'GC' (got %r); other localisation types are not "
                         "defined by the reference (observation/observation.py:82)" % (loc,))

    def _obs_arrays(self, loc_mode):
        """Structure-of-arrays view of self.obs with the forward-operator bookkeeping per ob."""
        st = self.prior
        obs = self.obs
        varnames = st.vars()
        nt, ny, nx = st.ntimes(), st.ny(), st.nx()
        vidx = {v: i for i, v in enumerate(varnames)}
        try:
            var = np.array([vidx[ob.obtype] for ob in obs], dtype=np.int64)
        except KeyError as e:
            raise KeyError('observation type %s is not a state variable %r' % (e, varnames))
        times = np.array([np.datetime64(ob.time) for ob in obs]).astype('datetime64[ns]')
        tlo, thi, wlo, whi, outside = engine.time_weights(st['validtime'].values, times)
        if outside.any():
            print("Interpolation is outside of time range in state!")
            raise ObTimeOutsideState("observation %d (time %s) is outside the state's valid times"
                                     % (int(np.argmax(outside)), obs[int(np.argmax(outside))].time))
        assim = np.array([1 if ob.assimilate_this else 0 for ob in obs], dtype=np.uint8)
        hw = np.ones(len(obs))
        if loc_mode == engine.LOC_GC:
            for k, ob in enumerate(obs):
                if ob.assimilate_this:
                    hw[k] = abs(ob.localize_radius)      # TypeError on None, as observation.py:120
        return engine.ObsArrays(
            value=np.array([ob.value for ob in obs], dtype=np.float64),
            error=np.array([ob.error for ob in obs], dtype=np.float64),
            lat=np.array([ob.lat for ob in obs], dtype=np.float64),
            lon=np.array([ob.lon for ob in obs], dtype=np.float64),
            halfwidth=hw, assimilate=assim,
            row0=(var * nt + tlo) * (ny * nx), row1=(var * nt + thi) * (ny * nx), tw0=wlo, tw1=whi)

    def _inflation_factors(self):
        """Per-level (variable x time) multiplicative factors, or None."""
        if self.inflation is None:
            return None
        st = self.prior
        varnames = st.vars()
        nt = st.ntimes()
        fac = np.ones((len(varnames), nt))
        if isinstance(self.inflation, float):
            fac[:] = self.inflation                                   # assimilation.py:62-69
        elif isinstance(self.inflation, dict):
            for k, v in self.inflation.items():                       # assimilation.py:82-114
                if k in ['validtime', 'lat', 'lon', 'x', 'y']:
                    raise NotImplementedError('per-dimension inflation arrays need xarray broadcasting and '
                                              'are not ported (assimilation.py:83-100)')
                assert isinstance(v, float)
                if k not in varnames:
                    print("Unable to find variable {:s} to inflate.  Skipping...".format(k))
                    continue
                fac[varnames.index(k), :] = v
        else:
            raise NotImplementedError('inflation from a netCDF file name needs xarray (assimilation.py:71-79)')
        return fac.ravel()

    def _device(self):
        import torch
        _lib.require_device()
        return torch.device('cuda', torch.cuda.current_device())

    # ---- reference methods --------------------------------------------------------------
    def compute_ob_priors(self):
        """Prior means [nobs] and perturbations [nobs, nmems] of every ob (assimilation.py:36-49)."""
        import torch
        dev = self._device()
        obs = self._obs_arrays(engine.LOC_NONE)
        X = torch.from_numpy(np.ascontiguousarray(self.prior.to_vect())).to(dev)
        grid = self.prior._grid_tables()
        Y, nex = engine.ob_priors(X, grid, obs, engine._sfx(X.dtype))
        self._check_exact(int(nex.item()))
        Ym = torch.empty(obs.nobs, dtype=X.dtype, device=dev)
        sfx = engine._sfx(X.dtype)
        _lib.call('exb_split_mean_pert_' + sfx, _lib.ptr(Y), _lib.ptr(Ym), obs.nobs, X.shape[1], _lib.stream_ptr())
        return Ym.cpu().numpy().astype(np.float64), Y.cpu().numpy().astype(np.float64)

    @staticmethod
    def _check_exact(n_exact):
        if n_exact > 0 and _pkg.EXACT_MATCH_POLICY == 'raise':
            raise IndexError('%d observation(s) lie within 1 km of a grid point: the reference raises here '
                             '(state/ensemble.py:195-196); set efa_xray_b200.EXACT_MATCH_POLICY = "nearest" '
                             'to use the nearest point instead' % n_exact)

    def inflate_state(self):
        """Inflate self.prior in place about its ensemble mean (assimilation.py:52-118)."""
        import ctypes as C
        import torch
        if self.is_inflated:
            print("State already inflated.  Skipping additional inflation.")
            return
        fac = self._inflation_factors()
        dev = self._device()
        X = torch.from_numpy(np.ascontiguousarray(self.prior.to_vect())).to(dev)
        nlev = fac.shape[0]
        sfx = engine._sfx(X.dtype)
        _lib.call('exb_inflate_' + sfx, _lib.ptr(X), X.shape[0], X.shape[1], fac.ctypes.data_as(C.c_void_p), nlev,
                  X.shape[0] // nlev, _lib.stream_ptr())
        self.prior.from_vect(X.cpu().numpy())
        self.is_inflated = True

    def format_prior_state(self):
        """(xbm [Nstate+Nobs], Xbp [Nstate+Nobs, Nens]): ensemble mean and perturbations of the state with
        the ob priors appended as extra rows (assimilation.py:120-154)."""
        import torch
        if self.inflation is not None:
            if self.verbose: print("Inflating Prior State")
            self.inflate_state()
        if self.verbose: print("Computing observation priors")
        obmeans, obperts = self.compute_ob_priors()
        if self.verbose: print("Converting state to vector")
        dev = self._device()
        X = torch.from_numpy(np.ascontiguousarray(self.prior.to_vect())).to(dev)
        xm = torch.empty(X.shape[0], dtype=X.dtype, device=dev)
        _lib.call('exb_split_mean_pert_' + engine._sfx(X.dtype), _lib.ptr(X), _lib.ptr(xm), X.shape[0], X.shape[1],
                  _lib.stream_ptr())
        xbm = np.hstack((xm.cpu().numpy(), obmeans))
        Xbp = np.vstack((X.cpu().numpy(), obperts))
        return xbm, Xbp

    def format_posterior_state(self, xam, Xap):
        """Posterior state object from analysis mean and perturbations (assimilation.py:157-171)."""
        if self.verbose: print("Formatting posterior")
        post_state = deepcopy(self.prior)
        Nstate = self.prior.nstate()
        post = (xam[:, None] + Xap)[:Nstate]
        post_state.from_vect(post)
        return post_state, self.obs


def update(prior_state, obs, inflate=None, loc=False, nproc=1, verbose=False):
    """Module-level entry point (assimilation.py:176-222).  In the reference this calls an undefined
    enkf_update; here it runs the EnSRF.  nproc is accepted for signature compatibility: the GPU path does
    not fork worker processes."""
    from .ensrf import EnSRF
    return EnSRF(prior_state, obs, nproc=nproc, inflation=inflate, verbose=verbose, loc=loc).update()


def randomize_obs_order(obs, seed=None):
    """Shuffle a list of observations in place and return it: the demo notebook assimilates in random order
    (efa_demo.ipynb cell 11, lines 44-46).  With localisation the analysis depends on the serial order, so the
    order used is part of the experiment; pass a seed to make it reproducible."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(obs))
    obs[:] = [obs[i] for i in perm]
    return obs
