"""CPU oracle: numpy restatement of efa_xray's serial EnSRF analysis step.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this module.  The product package
(efa_xray_b200/) never does; it fails loudly when its CUDA library is missing.

Pinned: tests/test_oracle_golden.py checks every function here against golden vectors produced by
running the UNMODIFIED reference package (tests/golden/make_golden.py, under a stand-in for the
uninstallable xarray).  The reference itself ships no tests or golden vectors.

The restatement works on plain arrays instead of xarray objects, but keeps the reference's dense
per-observation arithmetic (one-hot row select through a dot product, full-grid haversine, a
Python-level haversine call per observation pair, a materialised outer product), because that IS the
reference's CPU cost and this module is also the timed CPU baseline.  Citations are file:line under
/root/reference/efa_xray/.

Reference quirks that are reproduced on purpose:
  1. varye = np.var(ye) is /N (ensrf.py:69) while kcov is /(N-1) (ensrf.py:95).
  2. the forward operator is inverse-distance weighting over the 4 points nearest under the
     pseudo-metric hypot(d sin(lat), d cos(lon)) (state/ensemble.py:160-165), not bilinear.
  3. an ob within 1 km of a selected grid point raises IndexError (state/ensemble.py:195-196).
  4. time-interpolation weights are swapped (state/ensemble.py:218-224).
  5. un-flagged obs are skipped but still carried as obs-space rows (ensrf.py:74-76).
Ties in the nearest-point argsort (state/ensemble.py:165) are implementation-defined in numpy; this
oracle uses a stable sort, i.e. lowest flat index first.
"""
import numpy as np

R_EARTH = 6371.0   # state/ensemble.py:244, :259 ; observation/observation.py:138


class Ob(object):
    """Attribute record, observation/observation.py:17-37."""

    def __init__(self, value=None, obtype=None, time=None, error=None, lat=None, lon=None, vert=None,
                 prior_mean=None, post_mean=None, prior_var=None, post_var=None,
                 assimilate_this=False, description=None, localize_radius=None):
        self.value = value
        self.obtype = obtype
        self.time = time
        self.error = error
        self.lat = lat
        self.lon = lon
        self.vert = vert
        self.prior_mean = prior_mean
        self.post_mean = post_mean
        self.prior_var = prior_var
        self.post_var = post_var
        self.assimilate_this = assimilate_this
        self.assimilated = False
        self.description = description
        self.localize_radius = localize_radius


class State(object):
    """Plain-array stand-in for EnsembleState (state/ensemble.py:15): fields[name] has dims
    (validtime, y, x, mem); lat/lon are 2-D (y, x) -- or 1-D (x), the reference's second branch --; times is
    datetime64."""

    def __init__(self, fields, varnames, lat2d, lon2d, times):
        self.fields = {k: np.array(fields[k], dtype=np.float64, copy=True) for k in varnames}
        self.varnames = list(varnames)
        self.lat = np.asarray(lat2d, dtype=np.float64)
        self.lon = np.asarray(lon2d, dtype=np.float64)
        self.times = np.asarray(times).astype('datetime64[ns]')

    @classmethod
    def from_case(cls, case):
        return cls(case.fields, case.varnames, case.lat2d, case.lon2d, case.times)

    # state/ensemble.py:40-56
    def nmems(self):
        return self.fields[self.varnames[0]].shape[-1]

    def shape(self):
        return (len(self.varnames),) + self.fields[self.varnames[0]].shape

    def nstate(self):
        s = self.shape()
        return s[0] * s[1] * s[2] * s[3]

    def to_vect(self):
        """state/ensemble.py:110-114: [Nstate, Nens], row order var -> time -> y -> x."""
        arr = np.stack([self.fields[v] for v in self.varnames], axis=0)
        return np.reshape(arr, (self.nstate(), self.nmems()))

    def from_vect(self, instate):
        """state/ensemble.py:116-121."""
        instate = np.reshape(instate, self.shape())
        for i, v in enumerate(self.varnames):
            self.fields[v] = np.array(instate[i])

    def copy(self):
        return State(self.fields, self.varnames, self.lat, self.lon, self.times)


def obs_from_case(case):
    import datetime as _dt
    epoch = np.datetime64('1970-01-01T00:00:00', 's')
    obs = []
    for k in range(case.nobs):
        secs = int((case.ob_time[k] - epoch) / np.timedelta64(1, 's'))
        obs.append(Ob(value=float(case.ob_value[k]), obtype=case.varnames[int(case.ob_var[k])],
                      time=_dt.datetime(1970, 1, 1) + _dt.timedelta(seconds=secs),
                      error=float(case.ob_error[k]), lat=float(case.ob_lat[k]), lon=float(case.ob_lon[k]),
                      assimilate_this=bool(case.ob_assimilate[k]),
                      localize_radius=float(case.ob_halfwidth[k])))
    return obs


# ------------------------------------------------------------------------------------------
# distances and localisation
# ------------------------------------------------------------------------------------------
def haversine(loc1, loc2):
    """observation/observation.py:135-146 and state/ensemble.py:241-252 (identical bodies)."""
    lat1 = np.radians(loc1[0])
    lat2 = np.radians(loc2[0])
    dlat = lat2 - lat1
    dlon = np.radians(loc2[1] - loc1[1])
    a = np.sin(dlat / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin(dlon / 2) ** 2
    c = 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))
    return R_EARTH * c


def distance_to_point(state, lat, lon):
    """state/ensemble.py:254-267: haversine from every grid point to (lat, lon)."""
    lat = np.radians(lat)
    lon = np.radians(lon)
    dlat = lat - np.radians(state.lat)
    dlon = lon - np.radians(state.lon)
    a = np.sin(dlat / 2) ** 2 + np.cos(lat) * np.cos(np.radians(state.lat)) * np.sin(dlon / 2) ** 2
    c = 2 * np.arctan2(np.sqrt(a), np.sqrt(1.0 - a))
    return R_EARTH * c


def gaspari_cohn(distances, halfwidth):
    """observation/observation.py:117-130."""
    r = np.divide(distances, abs(halfwidth))
    weights = np.zeros(r.shape)
    with np.errstate(divide='ignore', invalid='ignore'):   # both branches are evaluated everywhere
        inner = ((((-0.25 * r + 0.5) * r + 0.625) * r - 5.0 / 3.0) * r ** 2 + 1.0)
        outer = (((((r / 12.0 - 0.5) * r + 0.625) * r + 5.0 / 3.0) * r - 5.0) * r + 4.0 - 2.0 / (3.0 * r))
    weights[r <= 1.0] = inner[r <= 1.0]
    m = (r > 1.0) & (r < 2.0)
    weights[m] = outer[m]
    return weights


def localize(ob, target, type='GC'):
    """observation/observation.py:59-87.  `target` is a State or a list of obs."""
    halfwidth = ob.localize_radius
    if isinstance(target, State):
        distances = distance_to_point(target, ob.lat, ob.lon)
    else:
        ourloc = (ob.lat, ob.lon)
        other_lats = [o.lat for o in target]
        other_lons = [o.lon for o in target]
        distances = np.array([haversine(ourloc, s) for s in zip(other_lats, other_lons)])
    if halfwidth is None:
        localization = np.ones(distances.shape)
    if type == 'GC':
        localization = gaspari_cohn(distances, halfwidth)   # abs(None) -> TypeError, as the reference
    return localization


# ------------------------------------------------------------------------------------------
# forward operator
# ------------------------------------------------------------------------------------------
def nearest_points(state, lat, lon, npt=1):
    """state/ensemble.py:152-168."""
    dist = np.hypot(np.sin(np.radians(state.lat)) - np.sin(np.radians(lat)),
                    np.cos(np.radians(state.lon)) - np.cos(np.radians(lon)))
    nearest_raw = dist.argsort(axis=None, kind='stable')[:npt]
    return np.unravel_index(nearest_raw, state.lat.shape)


def space_weights(state, lat, lon):
    """state/ensemble.py:179-200: 4 points, haversine, 1/d weights.  With 1-D lat/lon (:185-192) the 4 point indices
    are used for BOTH y and x, and everything carries a leading axis of length 1 (closen is a 1-tuple)."""
    if len(state.lat.shape) == 2:
        closey, closex = nearest_points(state, lat, lon, npt=4)
        distances = np.array([haversine((state.lat[y, x], state.lon[y, x]), (lat, lon))
                              for y, x in zip(list(closey), list(closex))])
    else:
        closen = nearest_points(state, lat, lon, npt=4)
        closey = closen
        closex = closen
        distances = np.array([haversine((state.lat[n], state.lon[n]), (lat, lon)) for n in list(closen)])
    spaceweights = np.zeros(distances.shape)
    if (distances < 1.0).sum() > 0:
        spaceweights[:, distances.argmin()] = 1       # IndexError, as the reference (trap 3)
    else:
        spaceweights = 1.0 / distances
        spaceweights /= spaceweights.sum()
    return closey, closex, spaceweights


def time_weights(state, time):
    """state/ensemble.py:202-224.  Returns None outside the valid range."""
    time64 = np.datetime64(time)
    valids = state.times
    timeweights = np.zeros(valids.shape)
    if (time64 < valids[0]) or (time64 > valids[-1]):
        return None
    lastdex = (valids >= time64).argmax()
    if valids[lastdex] == time64:
        timeweights[lastdex] = 1
    else:
        diff = (valids[lastdex] - valids[lastdex - 1])
        totsec = np.abs(diff / np.timedelta64(1, 's'))
        thisdiff = time64 - valids[lastdex]
        thissec = np.abs(thisdiff / np.timedelta64(1, 's'))
        timeweights[lastdex] = float(thissec) / totsec            # swapped on purpose (trap 4)
        timeweights[lastdex - 1] = 1.0 - (float(thissec) / totsec)
    return timeweights


def interpolate(state, var, time, lat, lon):
    """state/ensemble.py:170-239 -> [Nens]."""
    closey, closex, spaceweights = space_weights(state, lat, lon)
    timeweights = time_weights(state, time)
    if timeweights is None:
        return None
    interp = state.fields[var][:, closey, closex, :]                 # [nt, 4, Nens]   (1-D lat/lon: [nt, 1, 4, Nens])
    if len(interp.shape) == 3:                                       # state/ensemble.py:229-237
        interp = (timeweights[:, None, None] * interp).sum(axis=0)  # [4, Nens]
    else:
        interp = (timeweights[:, None, None, None] * interp).sum(axis=0)
    if len(interp.shape) == 3:
        interp = (spaceweights[:, :, None] * interp).sum(axis=1)    # [1, Nens]
    else:
        interp = (spaceweights[:, None] * interp).sum(axis=0)       # [Nens]
    return interp


def estimate(ob, state):
    """observation/observation.py:40-50."""
    return interpolate(state, ob.obtype, ob.time, ob.lat, ob.lon)


def compute_ob_priors(state, obs):
    """assimilation/assimilation.py:36-49."""
    nobs = len(obs)
    nmems = state.nmems()
    means = np.zeros(nobs)
    perts = np.zeros((nobs, nmems))
    for obnum, ob in enumerate(obs):
        ye = estimate(ob, state)
        means[obnum] = ye.mean()
        perts[obnum, :] = ye - ye.mean()
    return means, perts


# ------------------------------------------------------------------------------------------
# inflation (float, per-variable dict and per-dimension paths), assimilation/assimilation.py:52-118
# ------------------------------------------------------------------------------------------
def inflate_state(state, inflation):
    """Restates inflate_state.  Float (assimilation.py:62-69) and per-variable entries (:101-114) mutate `state`
    in place, as the reference does through `variables[v][:] = ...`.  Per-dimension entries ('validtime', 'x',
    'y'; :83-100) multiply the perturbations of EVERY variable by the array broadcast along that dimension --
    the reference rebinds `self.prior` to the new Dataset there, leaving the caller's object alone, so this
    function RETURNS the state the rest of the update must use (a copy in that case)."""
    if isinstance(inflation, float):
        for v in state.varnames:
            mean = state.fields[v].mean(axis=-1)
            perts = state.fields[v] - mean[..., None]
            state.fields[v][:] = perts * inflation + mean[..., None]
        return state
    for k, v in inflation.items():                                   # dict order, as the reference iterates
        if k in ['validtime', 'lat', 'lon', 'x', 'y']:
            v = np.asarray(v, dtype=np.float64)
            axis = {'validtime': 0, 'y': 1, 'x': 2}[k]                 # fields are [nt, ny, nx, nmem]
            assert v.shape[0] == next(iter(state.fields.values())).shape[axis]   # assimilation.py:87-88
            shape = [1, 1, 1, 1]
            shape[axis] = v.shape[0]
            new = state.copy()                                        # `self.prior = perts * infl + mean` (:96)
            for name in new.varnames:
                mean = state.fields[name].mean(axis=-1)
                perts = state.fields[name] - mean[..., None]
                new.fields[name] = perts * v.reshape(shape) + mean[..., None]
            state = new
        else:
            if k not in state.varnames:                              # "Unable to find variable ... Skipping" (:105-107)
                continue
            mean = state.fields[k].mean(axis=-1)
            perts = state.fields[k] - mean[..., None]
            state.fields[k][:] = perts * v + mean[..., None]
    return state


def format_prior_state(state, obs, inflation=None):
    """assimilation/assimilation.py:120-154."""
    if inflation is not None:
        state = inflate_state(state, inflation)
    obmeans, obperts = compute_ob_priors(state, obs)
    prior = state.to_vect()
    xbm = prior.mean(axis=1)
    Xbp = prior - xbm[:, None]
    xbm = np.hstack((xbm, obmeans))
    Xbp = np.vstack((Xbp, obperts))
    return xbm, Xbp, state


# ------------------------------------------------------------------------------------------
# the serial loop
# ------------------------------------------------------------------------------------------
def ensrf_loop(state, obs, xam, Xap, loc='GC', max_obs=None, timer=None):
    """assimilation/ensrf.py:37-40 and :50-149.  Mutates obs diagnostics; returns (xam, Xap).
    `max_obs` stops after that many loop iterations (bounded CPU-baseline samples)."""
    state_shape = state.shape()[:-1]
    dum_localize = np.ones(state_shape)
    Nstate = state.nstate()
    Nens = state.nmems()
    for obnum, ob in enumerate(obs):
        if max_obs is not None and obnum >= max_obs:
            break
        xbm = xam
        Xbp = Xap
        H = np.zeros(xam.shape)
        H[Nstate + obnum] = 1.0
        mye = np.dot(H, xbm)
        ye = np.dot(H, Xbp)
        ob.prior_mean = mye
        varye = np.var(ye)                                       # ddof = 0 (trap 1)
        ob.prior_var = varye
        if not ob.assimilate_this:
            ob.assimilated = False
            continue
        obs_err = ob.error
        innov = ob.value - mye
        kdenom = (varye + obs_err)
        kcov = np.dot(Xbp, np.transpose(ye)) / (Nens - 1)        # ddof = 1 (trap 1)
        if loc not in [None, False]:
            state_localize = localize(ob, state, type=loc)
            if len(state_localize.shape) == 2:
                state_localize = (state_localize[None, None, :, :] * dum_localize).flatten()
            else:
                state_localize = (state_localize[None, None, None, :] * dum_localize).flatten()
            obs_localize = localize(ob, obs, type=loc)
            state_localize = np.hstack((state_localize, obs_localize))
            kcov = np.multiply(state_localize, kcov)
        kmat = np.divide(kcov, kdenom)
        xam = xbm + np.multiply(kmat, innov)
        beta = 1. / (1. + np.sqrt(obs_err / (varye + obs_err)))
        kmat = np.multiply(beta, kmat)
        ye = np.array(ye)[np.newaxis]
        kmat = np.array(kmat)[np.newaxis]
        Xap = Xbp - np.dot(kmat.T, ye)
        post_ye = np.dot(H, xam)
        post_var = np.var(np.dot(H, Xap))
        ob.post_mean = post_ye
        ob.post_var = post_var
        ob.assimilated = True
    return xam, Xap


def ensrf_update(state, obs, loc='GC', inflation=None):
    """EnSRF(state, obs, inflation=..., loc=...).update(), assimilation/ensrf.py:33-151 with
    format_posterior_state (assimilation/assimilation.py:157-171).  Returns (post_state, obs);
    `state` is modified only by inflation, as in the reference."""
    xam, Xap, state = format_prior_state(state, obs, inflation)     # `state` is self.prior from here on
    xam, Xap = ensrf_loop(state, obs, xam, Xap, loc=loc)
    post_state = state.copy()
    Nstate = state.nstate()
    post = (xam[:, None] + Xap)[:Nstate]
    post_state.from_vect(post)
    return post_state, obs


# ------------------------------------------------------------------------------------------
# fast equivalent used ONLY to check large GPU runs (not the timed baseline)
# ------------------------------------------------------------------------------------------
def stencils_regular(case_lat2d, case_lon2d, ob_lat, ob_lon, chunk=256):
    """Vectorised restatement of nearest_points + the 1/d weights for many obs at once
    (state/ensemble.py:152-200).  Same arithmetic per ob; only the Python loop is batched."""
    sl = np.sin(np.radians(case_lat2d)).ravel()
    cl = np.cos(np.radians(case_lon2d)).ravel()
    latf = case_lat2d.ravel()
    lonf = case_lon2d.ravel()
    nobs = len(ob_lat)
    idx = np.empty((nobs, 4), dtype=np.int64)
    w = np.empty((nobs, 4))
    for c0 in range(0, nobs, chunk):
        c1 = min(nobs, c0 + chunk)
        d = np.hypot(sl[None, :] - np.sin(np.radians(ob_lat[c0:c1]))[:, None],
                     cl[None, :] - np.cos(np.radians(ob_lon[c0:c1]))[:, None])
        part = np.argpartition(d, 8, axis=1)[:, :9]
        dd = np.take_along_axis(d, part, axis=1)
        # stable order among the candidates: by (distance, flat index)
        order = np.lexsort((part, dd), axis=1)[:, :4]
        sel = np.take_along_axis(part, order, axis=1)
        idx[c0:c1] = sel
        for j in range(4):
            dist = haversine((latf[sel[:, j]], lonf[sel[:, j]]), (ob_lat[c0:c1], ob_lon[c0:c1]))
            w[c0:c1, j] = dist
    if (w < 1.0).any():
        raise IndexError('ob within 1 km of a selected grid point (state/ensemble.py:195-196)')
    w = 1.0 / w
    w /= w.sum(axis=1, keepdims=True)
    return idx, w


def obs_space_solve(ymean, ypert, ob_value, ob_error, ob_halfwidth, ob_lat, ob_lon, ob_assim, loc='GC'):
    """The obs-space rows evolved alone (they are a closed subsystem of ensrf.py:50-149, SURVEY.md
    section 0).  Vectorised over rows per ob; returns the per-ob records a state sweep needs."""
    nobs, nens = ypert.shape
    ym = ymean.copy()
    yp = ypert.copy()
    rec = dict(ye=np.zeros((nobs, nens)), prior_mean=np.zeros(nobs), prior_var=np.zeros(nobs),
               post_mean=np.full(nobs, np.nan), post_var=np.full(nobs, np.nan),
               innov=np.zeros(nobs), kdenom=np.ones(nobs), beta=np.zeros(nobs))
    for k in range(nobs):
        ye = yp[k].copy()
        mye = ym[k]
        varye = np.var(ye)
        rec['ye'][k] = ye
        rec['prior_mean'][k] = mye
        rec['prior_var'][k] = varye
        if not ob_assim[k]:
            continue
        innov = ob_value[k] - mye
        kdenom = varye + ob_error[k]
        kcov = yp[k:] @ ye / (nens - 1)
        if loc not in [None, False]:
            d = haversine((ob_lat[k], ob_lon[k]), (ob_lat[k:], ob_lon[k:]))
            kcov = gaspari_cohn(d, ob_halfwidth[k]) * kcov
        kmat = kcov / kdenom
        ym[k:] = ym[k:] + kmat * innov
        beta = 1. / (1. + np.sqrt(ob_error[k] / (varye + ob_error[k])))
        yp[k:] = yp[k:] - np.outer(beta * kmat, ye)
        rec['post_mean'][k] = ym[k]
        rec['post_var'][k] = np.var(yp[k])
        rec['innov'][k] = innov
        rec['kdenom'][k] = kdenom
        rec['beta'][k] = beta
    return rec
