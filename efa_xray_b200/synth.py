"""Synthetic EnSRF cases ("Synthetic inputs").

Pure numpy input generation -- no assimilation arithmetic lives here.  Used by the
tests, by tests/golden/make_golden.py (which feeds the same arrays to the real
reference) and by bench.py.

Grid      regular global lat-lon, lat = linspace(-90, 90, ny), lon = arange(nx)*360/nx,
          2-D via meshgrid (the reference wants lat(y,x) / lon(y,x), state/ensemble.py:178).
Ensemble  smooth base field 288 - 40 sin^2(lat) (+10 K per variable index, +0.5 K per time
          index) + per-member low-wavenumber waves + white noise (sigma 0.3).
Obs       uniform on the sphere (lat = asin(U(-0.98, 0.98)), lon = U(0, 360)), rejecting any ob
          within `min_grid_km` of a nearby grid point (the reference raises IndexError when an ob
          is within 1 km of one of its 4 selected points, state/ensemble.py:195-196).
          value = truth wave field at the ob + N(0, sqrt(R)); error R is a variance.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

EARTH_RADIUS_KM = 6371.0

# (zonal wavenumber, meridional wavenumber) of the low-wavenumber member perturbations
_WAVES = ((1, 1), (2, 1), (3, 2), (4, 3), (2, 3), (5, 2))


@dataclass
class Case:
    """Plain-array description of one synthetic analysis problem."""
    lat2d: np.ndarray            # [ny, nx] degrees
    lon2d: np.ndarray            # [ny, nx] degrees in [0, 360)
    times: np.ndarray            # [nt] datetime64[s]
    varnames: list               # nvars names, order = state-vector order
    fields: dict                 # name -> [nt, ny, nx, nmem] float64 (mem last, as the reference)
    ob_value: np.ndarray         # [nobs]
    ob_lat: np.ndarray           # [nobs] degrees
    ob_lon: np.ndarray           # [nobs] degrees
    ob_time: np.ndarray          # [nobs] datetime64[s]
    ob_var: np.ndarray           # [nobs] int index into varnames
    ob_error: np.ndarray         # [nobs] error VARIANCE
    ob_halfwidth: np.ndarray     # [nobs] Gaspari-Cohn half-width c in km (support is 2c)
    ob_assimilate: np.ndarray    # [nobs] bool
    meta: dict = field(default_factory=dict)

    @property
    def nobs(self):
        return int(self.ob_value.shape[0])

    @property
    def nmem(self):
        return int(next(iter(self.fields.values())).shape[-1])

    def to_vect(self):
        """[Nstate, Nens] in the reference's row order var -> time -> y -> x (ensemble.py:110-114)."""
        arr = np.stack([self.fields[v] for v in self.varnames], axis=0)
        return arr.reshape(-1, arr.shape[-1])


def regular_grid(ny, nx):
    lat = np.linspace(-90.0, 90.0, ny)
    lon = np.arange(nx) * (360.0 / nx)
    lon2d, lat2d = np.meshgrid(lon, lat)
    return np.ascontiguousarray(lat2d), np.ascontiguousarray(lon2d)


def _haversine_km(lat1, lon1, lat2, lon2):
    p1, p2 = np.radians(lat1), np.radians(lat2)
    dlat = p2 - p1
    dlon = np.radians(lon2 - lon1)
    a = np.sin(dlat / 2) ** 2 + np.cos(p1) * np.cos(p2) * np.sin(dlon / 2) ** 2
    return EARTH_RADIUS_KM * 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))


def _wave_basis(lat_deg, lon_deg, phases):
    """[..., W] values of the W wave modes at the given points."""
    lam = np.radians(lon_deg)[..., None]
    phi = np.radians(lat_deg)[..., None]
    k = np.array([w[0] for w in _WAVES], dtype=np.float64)
    l = np.array([w[1] for w in _WAVES], dtype=np.float64)
    return np.cos(k * lam + phases[:, 0]) * np.cos(l * phi + phases[:, 1]) * np.cos(phi)


def _base(lat_deg):
    return 288.0 - 40.0 * np.sin(np.radians(lat_deg)) ** 2


def draw_obs_locations(rng, nobs, ny, nx, min_grid_km=1.2, avoid_mirror_ties=False):
    """Uniform-on-sphere points that are not within `min_grid_km` of any nearby grid point.

    avoid_mirror_ties: also reject points within 2.5 grid columns of the 0 and 180 degree meridians.
    The reference's pseudo-metric cannot tell longitude L from 360-L (their cosines are bit-identical),
    and next to those two self-mirror columns the exact tie falls on the 4th/5th-nearest boundary, so
    which point the reference picks depends on numpy's unstable argsort (state/ensemble.py:165).
    Golden cases stay out of that zone; this package and the oracle break ties by lowest flat index."""
    dlat = 180.0 / (ny - 1)
    dlon = 360.0 / nx
    lats = np.empty(0)
    lons = np.empty(0)
    while lats.size < nobs:
        n = int((nobs - lats.size) * 1.1) + 16
        la = np.degrees(np.arcsin(rng.uniform(-0.98, 0.98, n)))
        lo = rng.uniform(0.0, 360.0, n)
        jy = np.rint((la + 90.0) / dlat).astype(np.int64)
        jx = np.rint(lo / dlon).astype(np.int64)
        dmin = np.full(n, np.inf)
        for oy in (-1, 0, 1):
            gy = np.clip(jy + oy, 0, ny - 1)
            for ox in (-1, 0, 1):
                gx = (jx + ox) % nx
                d = _haversine_km(-90.0 + gy * dlat, gx * dlon, la, lo)
                dmin = np.minimum(dmin, d)
        keep = dmin >= min_grid_km
        if avoid_mirror_ties:
            off = np.minimum(np.abs(((lo + 90.0) % 180.0) - 90.0), 180.0)   # distance to 0/180 meridian
            keep &= off > 2.5 * dlon
        lats = np.concatenate([lats, la[keep]])
        lons = np.concatenate([lons, lo[keep]])
    return lats[:nobs].copy(), lons[:nobs].copy()


def make_case(ny=181, nx=360, nmem=50, nvars=1, ntimes=1, nobs=500, cutoff_km=2000.0,
              seed=0, frac_skip=0.0, mixed_error=False, offtime=False, noise_sigma=0.3,
              mixed_radius=False, avoid_mirror_ties=False, dtype=np.float64, out=None):
    """Build a Case.  `cutoff_km` is the localisation support radius; half-width = cutoff/2.

    `out`, if given, is a preallocated [nvars, ntimes, ny, nx, nmem] array (e.g. pinned host
    memory) that receives the ensemble; the Case's fields are views into it.
    """
    rng = np.random.default_rng(seed)
    lat2d, lon2d = regular_grid(ny, nx)
    times = (np.datetime64('2020-01-01T00:00:00', 's') + np.arange(ntimes) * np.timedelta64(6 * 3600, 's'))
    varnames = ['var%d' % v for v in range(nvars)]
    if nvars == 1:
        varnames = ['t2m']

    phases = rng.uniform(0, 2 * np.pi, (len(_WAVES), 2))
    basis = _wave_basis(lat2d, lon2d, phases).reshape(ny * nx, len(_WAVES))   # [npts, W]
    base = _base(lat2d).reshape(ny * nx)
    if out is None:
        out = np.empty((nvars, ntimes, ny, nx, nmem), dtype=dtype)
    assert out.shape == (nvars, ntimes, ny, nx, nmem)
    amp_truth = rng.normal(0.0, 1.0, (nvars, ntimes, len(_WAVES)))
    for v in range(nvars):
        for t in range(ntimes):
            amp = rng.normal(0.0, 1.0, (len(_WAVES), nmem))                # member amplitudes
            fld = basis @ amp                                               # [npts, nmem]
            fld += (base + 10.0 * v + 0.5 * t)[:, None]
            # white noise in row blocks to bound temporaries
            blk = max(1, (1 << 22) // nmem)
            for r0 in range(0, ny * nx, blk):
                r1 = min(ny * nx, r0 + blk)
                fld[r0:r1] += noise_sigma * rng.standard_normal((r1 - r0, nmem))
            out[v, t] = fld.reshape(ny, nx, nmem)
    fields = {name: out[v] for v, name in enumerate(varnames)}

    ob_lat, ob_lon = draw_obs_locations(rng, nobs, ny, nx, avoid_mirror_ties=avoid_mirror_ties)
    ob_var = rng.integers(0, nvars, nobs)
    if offtime and ntimes > 1:
        tfrac = rng.uniform(0.0, ntimes - 1.0, nobs)
        secs = np.rint(tfrac * 6 * 3600).astype(np.int64)
        ob_time = times[0] + secs * np.timedelta64(1, 's')
        tpos = secs / (6.0 * 3600.0)
    else:
        tix = rng.integers(0, ntimes, nobs)
        ob_time = times[tix]
        tpos = tix.astype(np.float64)
    ob_error = rng.uniform(0.5, 2.0, nobs) if mixed_error else np.full(nobs, 1.0)
    halfwidth = np.full(nobs, 0.5 * cutoff_km)
    if mixed_radius:
        halfwidth = halfwidth * rng.choice([0.5, 1.0, 2.0], nobs)
    ob_basis = _wave_basis(ob_lat, ob_lon, phases)                             # [nobs, W]
    # truth amplitude linearly interpolated in time (only matters for offtime obs)
    t0 = np.clip(np.floor(tpos).astype(np.int64), 0, ntimes - 1)
    t1 = np.clip(t0 + 1, 0, ntimes - 1)
    w1 = tpos - t0
    a = amp_truth[ob_var, t0] * (1 - w1)[:, None] + amp_truth[ob_var, t1] * w1[:, None]
    truth = _base(ob_lat) + 10.0 * ob_var + 0.5 * tpos + (ob_basis * a).sum(axis=1)
    ob_value = truth + np.sqrt(ob_error) * rng.standard_normal(nobs)
    ob_assim = np.ones(nobs, dtype=bool)
    if frac_skip > 0:
        ob_assim = rng.uniform(0, 1, nobs) >= frac_skip
    return Case(lat2d=lat2d, lon2d=lon2d, times=times.astype('datetime64[s]'), varnames=varnames,
                fields=fields, ob_value=ob_value, ob_lat=ob_lat, ob_lon=ob_lon,
                ob_time=ob_time.astype('datetime64[s]'), ob_var=ob_var.astype(np.int64),
                ob_error=ob_error, ob_halfwidth=halfwidth, ob_assimilate=ob_assim,
                meta=dict(ny=ny, nx=nx, nmem=nmem, nvars=nvars, ntimes=ntimes, nobs=nobs,
                          cutoff_km=cutoff_km, seed=seed, frac_skip=frac_skip,
                          mixed_error=mixed_error, offtime=offtime, mixed_radius=mixed_radius,
                          avoid_mirror_ties=avoid_mirror_ties))


def make_case_1d(npts=24, nmem=8, nvars=1, ntimes=1, nobs=30, cutoff_km=4000.0, seed=0, frac_skip=0.0):
    """A state with 1-D lat(x)/lon(x): the reference's second coordinate branch (state/ensemble.py:185-192,
    assimilation/ensrf.py:110-111).  There the forward operator reads the state at (y, x) = (n, n) for a point index n
    and the localisation weight of row (y, x) depends on x only, so the state is [nt, npts, npts, nmem] over a list of
    npts scattered points.  Case.lat2d / lon2d hold the 1-D coordinate arrays."""
    rng = np.random.default_rng(seed)
    lat = np.degrees(np.arcsin(rng.uniform(-0.9, 0.9, npts)))
    lon = rng.uniform(0.0, 360.0, npts)
    times = (np.datetime64('2020-01-01T00:00:00', 's') + np.arange(ntimes) * np.timedelta64(6 * 3600, 's'))
    varnames = ['var%d' % v for v in range(nvars)] if nvars > 1 else ['t2m']
    phases = rng.uniform(0, 2 * np.pi, (len(_WAVES), 2))
    basis = _wave_basis(lat, lon, phases)                                     # [npts, W]
    out = np.empty((nvars, ntimes, npts, npts, nmem))
    amp_truth = rng.normal(0.0, 1.0, (nvars, ntimes, len(_WAVES)))
    for v in range(nvars):
        for t in range(ntimes):
            amp = rng.normal(0.0, 1.0, (len(_WAVES), nmem))
            fld = basis @ amp + (_base(lat) + 10.0 * v + 0.5 * t)[:, None]     # [npts(x), nmem]
            out[v, t] = fld[None, :, :] + 0.3 * rng.standard_normal((npts, npts, nmem))
    fields = {name: out[v] for v, name in enumerate(varnames)}
    # obs a few hundred km away from the points (never within 1 km of one)
    k = rng.integers(0, npts, nobs)
    ob_lat = np.clip(lat[k] + rng.uniform(1.0, 4.0, nobs) * rng.choice([-1.0, 1.0], nobs), -88.0, 88.0)
    ob_lon = (lon[k] + rng.uniform(1.0, 4.0, nobs) * rng.choice([-1.0, 1.0], nobs)) % 360.0
    ob_var = rng.integers(0, nvars, nobs)
    tix = rng.integers(0, ntimes, nobs)
    ob_basis = _wave_basis(ob_lat, ob_lon, phases)
    truth = _base(ob_lat) + 10.0 * ob_var + 0.5 * tix + (ob_basis * amp_truth[ob_var, tix]).sum(axis=1)
    ob_error = np.full(nobs, 1.0)
    ob_value = truth + rng.standard_normal(nobs)
    ob_assim = rng.uniform(0, 1, nobs) >= frac_skip if frac_skip > 0 else np.ones(nobs, dtype=bool)
    return Case(lat2d=lat, lon2d=lon, times=times.astype('datetime64[s]'), varnames=varnames, fields=fields,
                ob_value=ob_value, ob_lat=ob_lat, ob_lon=ob_lon, ob_time=times[tix].astype('datetime64[s]'),
                ob_var=ob_var.astype(np.int64), ob_error=ob_error, ob_halfwidth=np.full(nobs, 0.5 * cutoff_km),
                ob_assimilate=ob_assim,
                meta=dict(npts=npts, nmem=nmem, nvars=nvars, ntimes=ntimes, nobs=nobs, cutoff_km=cutoff_km, seed=seed,
                          frac_skip=frac_skip, one_d=True))


# BASELINE.json configs (sizes from SURVEY.md section 8 header)
CONFIGS = {
    'config1': dict(ny=181, nx=360, nmem=50, nvars=1, ntimes=1, nobs=500),
    'config2': dict(ny=361, nx=720, nmem=50, nvars=3, ntimes=4, nobs=5000),
    'config3': dict(ny=721, nx=1440, nmem=100, nvars=3, ntimes=1, nobs=100000),
    'config4': dict(ny=721, nx=1440, nmem=100, nvars=10, ntimes=1, nobs=100000),
}


def build_objects(case, state_cls, ob_cls):
    """Turn a Case into (state, [obs]) for any package that follows the reference API
    (EnsembleState.from_vardict, ensemble.py:25-36; Observation(...), observation.py:18-36)."""
    import datetime as _dt
    ny, nx = next(iter(case.fields.values())).shape[1:3]
    vardict = {name: (('validtime', 'y', 'x', 'mem'), case.fields[name]) for name in case.varnames}
    cdims = ('y', 'x') if case.lat2d.ndim == 2 else ('x',)         # 1-D lat(x)/lon(x): make_case_1d
    coorddict = {'validtime': case.times.astype('datetime64[ns]'),
                 'lat': (cdims, case.lat2d), 'lon': (cdims, case.lon2d),
                 'mem': np.arange(case.nmem) + 1, 'y': np.arange(ny), 'x': np.arange(nx)}
    state = state_cls.from_vardict(vardict, coorddict)
    obs = []
    epoch = np.datetime64('1970-01-01T00:00:00', 's')
    for k in range(case.nobs):
        secs = int((case.ob_time[k] - epoch) / np.timedelta64(1, 's'))
        t = _dt.datetime(1970, 1, 1) + _dt.timedelta(seconds=secs)
        obs.append(ob_cls(value=float(case.ob_value[k]), obtype=case.varnames[int(case.ob_var[k])],
                          time=t, error=float(case.ob_error[k]), lat=float(case.ob_lat[k]),
                          lon=float(case.ob_lon[k]), assimilate_this=bool(case.ob_assimilate[k]),
                          description='synthetic %d' % k,
                          localize_radius=float(case.ob_halfwidth[k])))
    return state, obs
