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
        st = self.prior
        nlev = st.nvars() * st.ntimes()

        if self.inflation is not None and not self.is_inflated:
            # the reference inflates the caller's state in place before anything else (assimilation.py:132-134)
            if self.verbose: print("Inflating Prior State")
            self.inflate_state()

        if self.verbose: print("Computing observation priors")
        obs = self._obs_arrays(loc_mode)
        host = np.array(st.to_vect(), order='C', copy=True)     # the prior itself stays untouched (ensrf.py:165)
        tdtype = {'f64': torch.float64, 'f32': torch.float32}[self.dtype]
        if self.verbose: print("Beginning observation loop")
        # host buffer in, analysis written back into it (upload, sweep and download overlap band by band)
        res = engine.analysis_host(host, nlev, None, None, obs, loc_mode, device=dev, dtype=tdtype,
                                   grid=st._grid_tables())
        self._check_exact(res.n_exact)
        self.last_result = res

        # per-ob diagnostics back onto the caller's objects (ensrf.py:66-76, :144-149)
        pm, pv = res.prior_mean.tolist(), res.prior_var.tolist()
        qm, qv = res.post_mean.tolist(), res.post_var.tolist()
        done = res.assimilated.tolist()
        for k, ob in enumerate(self.obs):
            ob.prior_mean = pm[k]
            ob.prior_var = pv[k]
            if done[k]:
                ob.post_mean = qm[k]
                ob.post_var = qv[k]
                ob.assimilated = True
            else:
                ob.assimilated = False

        if self.verbose: print("Formatting posterior")
        post_state = deepcopy(self.prior)
        post_state.from_vect(host)
        return post_state, self.obs
