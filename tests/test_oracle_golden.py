"""The CPU oracle against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA path
with the oracle and with the same golden vectors."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, golden_case
from efa_xray_b200.synth import make_case
from oracle import ensrf_oracle as O


def _diag(obs, attr):
    return np.array([np.nan if getattr(o, attr) is None else float(getattr(o, attr)) for o in obs])


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_full_update_matches_reference(name):
    g, p = load_golden(name)
    case = golden_case(p)
    st, obs = O.State.from_case(case), O.obs_from_case(case)
    prior = st.to_vect()
    np.testing.assert_allclose([prior.sum(), np.abs(prior).sum()], g['prior_checksum'], rtol=1e-13)
    ye = np.array([O.estimate(o, st) for o in obs])
    np.testing.assert_allclose(ye, g['ye'], rtol=1e-13)
    near = np.array([np.array(O.nearest_points(st, o.lat, o.lon, 4)) for o in obs])
    for i in range(len(obs)):       # order among exact mirror ties is sort-implementation defined
        assert set(map(tuple, near[i].T)) == set(map(tuple, g['nearest'][i].T))
    post, obs = O.ensrf_update(st, obs, loc=p['loc'], inflation=p['inflation'])
    np.testing.assert_allclose(post.to_vect(), g['post'], rtol=1e-12)
    # the caller's state afterwards: inflated in place by the float / per-variable forms, untouched by arrays
    after = st.to_vect()
    np.testing.assert_allclose([after.sum(), np.abs(after).sum()], g['prior_after_checksum'], rtol=1e-13)
    for attr in ('prior_mean', 'prior_var', 'post_mean', 'post_var'):
        np.testing.assert_allclose(_diag(obs, attr), g[attr], rtol=1e-11, equal_nan=True)
    assert np.array_equal([o.assimilated for o in obs], g['assimilated'])


def test_localization_vectors():
    g, p = load_golden('gc_small')
    case = golden_case(p)
    st, obs = O.State.from_case(case), O.obs_from_case(case)
    np.testing.assert_allclose(O.localize(obs[0], st), g['loc_state0'], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(O.localize(obs[0], obs), g['loc_obs0'], rtol=1e-13, atol=1e-16)


def test_free_functions():
    import os
    from conftest import GOLDEN
    f = np.load(os.path.join(GOLDEN, 'functions.npz'))
    np.testing.assert_allclose(O.gaspari_cohn(f['gc_d'], 1000.0), f['gc_w'], rtol=1e-14, atol=1e-17)
    np.testing.assert_allclose(O.gaspari_cohn(f['gc_d'], -1000.0), f['gc_w_neg'], rtol=1e-14, atol=1e-17)
    hv = np.array([O.haversine((q[0], q[1]), (q[2], q[3])) for q in f['hv_pairs']])
    np.testing.assert_allclose(hv, f['hv_km'], rtol=1e-14, atol=1e-9)
    assert str(f['exact_point_raises']) == 'IndexError'
    assert bool(f['outside_time_is_none'])


def test_known_answers():
    # Gaspari-Cohn: w(0)=1, w(c)=5/24, w(2c)=0, continuous at r=1, zero beyond 2c (observation.py:125-129)
    w = O.gaspari_cohn(np.array([0.0, 1000.0, 2000.0, 999.9999999, 1000.0000001, 3000.0]), 1000.0)
    assert w[0] == 1.0 and abs(w[1] - 5.0 / 24.0) < 1e-15 and w[2] == 0.0 and w[5] == 0.0
    assert abs(w[3] - w[4]) < 1e-9
    # haversine (0,0)->(0,90) is a quarter great circle
    assert abs(O.haversine((0, 0), (0, 90)) - 6371.0 * np.pi / 2) < 1e-9


def test_single_ob_closed_form():
    """One ob, no localisation: post_mean = mye + K innov with K = (N/(N-1)) var/(var+R) and
    post_var = var (1 - beta K)^2 (ensrf.py:69,95,119,135,141)."""
    case = make_case(ny=19, nx=36, nmem=12, nobs=1, seed=5)
    st, obs = O.State.from_case(case), O.obs_from_case(case)
    obs[0].assimilate_this = True
    ye = O.estimate(obs[0], st)
    N, var, R = 12, np.var(ye), obs[0].error
    K = (N / (N - 1.0)) * var / (var + R)
    beta = 1.0 / (1.0 + np.sqrt(R / (var + R)))
    O.ensrf_update(st, obs, loc=False)
    assert abs(obs[0].post_mean - (ye.mean() + K * (obs[0].value - ye.mean()))) < 1e-11
    assert abs(obs[0].post_var - var * (1 - beta * K) ** 2) < 1e-12


def test_skipped_ob_leaves_state_unchanged():
    case = make_case(ny=19, nx=36, nmem=6, nobs=3, seed=6)
    st, obs = O.State.from_case(case), O.obs_from_case(case)
    for o in obs:
        o.assimilate_this = False
    post, obs = O.ensrf_update(st, obs, loc='GC')
    np.testing.assert_allclose(post.to_vect(), st.to_vect(), rtol=1e-15)
    assert all(o.prior_mean is not None and o.prior_var is not None and not o.assimilated and o.post_mean is None
               for o in obs)


def test_obs_space_is_closed_and_order_matters():
    """The obs rows evolved alone reproduce the augmented run; permuting obs under
    localisation changes the analysis."""
    case = make_case(ny=25, nx=48, nmem=10, nobs=40, cutoff_km=4000.0, seed=7, frac_skip=0.1)
    st, obs = O.State.from_case(case), O.obs_from_case(case)
    ym, yp = O.compute_ob_priors(st, obs)
    rec = O.obs_space_solve(ym, yp, case.ob_value, case.ob_error, case.ob_halfwidth, case.ob_lat, case.ob_lon,
                            case.ob_assimilate)
    post, obs = O.ensrf_update(st, obs, loc='GC')
    np.testing.assert_allclose(rec['prior_mean'], [o.prior_mean for o in obs], rtol=1e-12)
    np.testing.assert_allclose(rec['prior_var'], [o.prior_var for o in obs], rtol=1e-11)
    done = case.ob_assimilate
    np.testing.assert_allclose(rec['post_var'][done], [o.post_var for o in obs if o.assimilated], rtol=1e-11)
    st2, obs2 = O.State.from_case(case), O.obs_from_case(case)
    post2, _ = O.ensrf_update(st2, obs2[::-1], loc='GC')
    assert np.abs(post2.to_vect() - post.to_vect()).max() > 1e-6


def test_vectorised_stencils_match_loop():
    case = make_case(ny=46, nx=90, nmem=4, nobs=50, seed=8)
    st = O.State.from_case(case)
    idx, w = O.stencils_regular(case.lat2d, case.lon2d, case.ob_lat, case.ob_lon)
    for k in range(case.nobs):
        cy, cx, sw = O.space_weights(st, case.ob_lat[k], case.ob_lon[k])
        assert set(idx[k]) == set(cy * 90 + cx)
        np.testing.assert_allclose(np.sort(w[k]), np.sort(sw), rtol=1e-12)
