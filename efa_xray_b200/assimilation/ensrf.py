"""EnSRF -- serial ensemble square-root filter update (efa_xray/assimilation/ensrf.py:8-151)."""
from __future__ import print_function

from copy import deepcopy

import numpy as np

from .assimilation import Assimilation
from .. import engine


class EnSRF(Assimilation):
    """EnSRF(state, obs, nproc=1, inflation=None, verbose=True, loc=False).update() -> (post_state, obs).

    Same contract as the reference (ensrf.py:28-33, assimilation.py:171): `obs` is the caller's list,
    mutated in place with prior_mean / prior_var / post_mean / post_var / assimilated (ensrf.py:66-70,
    :144-149); `post_state` is a new EnsembleState; the prior is only modified by inflation.
    Observations are assimilated strictly in list order.  `dtype` ('f64' or 'f32') selects the device
    arithmetic; results are returned in the state's own dtype.
    """

    def __init__(self, state, obs, nproc=1, inflation=None, verbose=True, loc=False, dtype='f64'):
        Assimilation.__init__(self, state, obs, nproc, inflation, verbose)
        self.loc = loc
        self.dtype = dtype
        self.last_result = None

    def update(self):
        import torch
        if self.verbose: print("Beginning update sequence")
        loc_mode = self._loc_mode(self.loc)
        dev = self._device()

        if self.inflation is not None and not self.is_inflated:
            # the reference inflates the prior before anything else (assimilation.py:132-134): the caller's state in
            # place, except that per-dimension arrays rebind self.prior to a new state (inflate_state)
            if self.verbose: print("Inflating Prior State")
            self.inflate_state()
        st = self.prior
        nlev = st.nvars() * st.ntimes()

        if self.verbose: print("Computing observation priors")
        obs = self._obs_arrays(loc_mode)
        tdtype = {'f64': torch.float64, 'f32': torch.float32}[self.dtype]
        if self.verbose: print("Beginning observation loop")
        # The prior is read where it lies (the state's contiguous, page-locked block: no to_vect copy, and the prior
        # itself stays untouched, assimilation.py:165) and the analysis is written into a fresh block that the
        # posterior state adopts (no deepcopy of the data, no from_vect copy).  Upload, sweep and download overlap
        # band by band inside engine.analysis_host.
        st._consolidate()
        prior_blk = st._block_view()
        if prior_blk is not None:
            from ..state.ensemble import BLOCK_POOL
            host = prior_blk.reshape(st.nstate(), st.nmems())
            out_blk, owner = BLOCK_POOL.alloc(prior_blk.shape, prior_blk.dtype)
            out = out_blk.reshape(host.shape)
        else:                                    # variables on different dims / dtypes: stacked copy
            host = np.array(st.to_vect(), order='C', copy=True)
            out_blk, owner, out = None, None, host
        res = engine.analysis_host(host, nlev, None, None, obs, loc_mode, device=dev, dtype=tdtype,
                                   grid=st._grid_tables(), out=out)
        self._check_exact(res.n_exact)
        self.last_result = res

        # per-ob diagnostics back onto the caller's objects (ensrf.py:66-76, :144-149)
        pm, pv = res.prior_mean.tolist(), res.prior_var.tolist()
        qm, qv = res.post_mean.tolist(), res.post_var.tolist()
        done = res.assimilated.tolist()
        for ob, a, b, c, d, f in zip(self.obs, pm, pv, qm, qv, done):
            ob.prior_mean = a
            ob.prior_var = b
            if f:
                ob.post_mean = c
                ob.post_var = d
                ob.assimilated = True
            else:
                ob.assimilated = False

        if self.verbose: print("Formatting posterior")
        if out_blk is not None:
            post_state = st._new_with_block(out_blk, owner)
        else:
            post_state = deepcopy(self.prior)
            post_state.from_vect(host)
        return post_state, self.obs
