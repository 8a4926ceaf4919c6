"""Host-side logic that needs no GPU: band schedules, ob-order helper, time weights, sharding helpers."""
import numpy as np
import pytest


def test_sweep_band_schedule_covers_rows_on_patch_boundaries(lib):
    from efa_xray_b200 import engine
    for nlev, ny, nx in ((3, 721, 1440), (1, 181, 360), (12, 361, 720), (10, 721, 1440), (3, 90, 1440), (2, 7, 9)):
        g = lib.exb_state_sweep_row_granularity(nlev, ny, nx)
        assert 1 <= g <= 12
        bands = engine.sweep_band_schedule(nlev, ny, nx)
        assert bands[0][0] == 0 and bands[-1][1] == ny
        for (a, b), (c, d) in zip(bands[:-1], bands[1:]):
            assert b == c and a < b
        for a, b in bands[:-1]:
            assert b % g == 0          # a patch is never split between two calls
        assert len(bands) <= 7
    # the last band is the short one (its download is what stays exposed)
    bands = engine.sweep_band_schedule(3, 721, 1440)
    assert bands[-1][1] - bands[-1][0] <= (bands[0][1] - bands[0][0])


def test_randomize_obs_order_is_a_seeded_permutation():
    from efa_xray.assimilation.assimilation import randomize_obs_order
    obs = list(range(50))
    a = randomize_obs_order(list(obs), seed=3)
    b = randomize_obs_order(list(obs), seed=3)
    c = randomize_obs_order(list(obs), seed=4)
    assert a == b and a != c and sorted(a) == obs and a != obs
    same = list(obs)
    assert randomize_obs_order(same, seed=1) is same          # in place


def test_time_weights_follow_the_reference_including_its_swap():
    """state/ensemble.py:202-224: an ob exactly on a valid time gets weight 1 there; between two times the
    reference attaches the LARGER weight to the FARTHER time (its weights are swapped); outside -> flagged."""
    from efa_xray_b200 import engine
    valid = np.array(['2020-01-01T00', '2020-01-01T06', '2020-01-01T12'], dtype='datetime64[s]')
    t = np.array(['2020-01-01T06', '2020-01-01T07:30', '2020-01-01T00', '2019-12-31T23', '2020-01-01T13'], dtype='datetime64[s]')
    lo, hi, wlo, whi, outside = engine.time_weights(valid, t)
    assert outside.tolist() == [False, False, False, True, True]
    assert (hi[0], whi[0], wlo[0]) == (1, 1.0, 0.0)
    assert (lo[1], hi[1]) == (1, 2)
    np.testing.assert_allclose([wlo[1], whi[1]], [0.25, 0.75])   # 1.5 h after 06: upper level (4.5 h away) gets 0.75
    assert (hi[2], whi[2]) == (0, 1.0)


def test_postprocess_module_is_reexported():
    import efa_xray.postprocess.postprocess as p
    from efa_xray_b200.postprocess.postprocess import COLUMNS
    assert callable(p.obs_assimilation_statistics)
    assert COLUMNS[:4] == ['validtime', 'flead', 'lat', 'lon'] and COLUMNS[-1] == 'post variance'


def test_localize_stencil_keeps_only_band_rows():
    from efa_xray_b200.sharding import localize_stencil, partition_bands, equal_bands
    nlev, ny, nx = 2, 10, 4
    idx = np.array([[0, 5 * nx + 1, ny * nx + 9 * nx + 3, 3 * nx]])
    w = np.array([[0.1, 0.2, 0.3, 0.4]])
    li, lw = localize_stencil(idx, w, nlev, ny, nx, 3, 6)
    assert lw.tolist() == [[0.0, 0.2, 0.0, 0.4]]
    assert li[0, 1] == (0 * 3 + 2) * nx + 1 and li[0, 3] == 0 * nx + 0
    assert equal_bands(10, 3) == [(0, 3), (3, 7), (7, 10)]
    assert partition_bands(np.ones(10), 2) == [(0, 5), (5, 10)]
