"""Host-side multi-GPU logic on CPU: latitude-band partition, stencil re-indexing and the three
collectives of efa_xray_b200.sharding, with world size 2 over gloo (no GPU, no compute kernels).

The per-band arithmetic is emulated with numpy here (the oracle's gather) only to check that the plumbing
moves the right rows to the right rank; the CUDA kernels are covered by tests/test_gpu_parity.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from efa_xray_b200 import sharding
from efa_xray_b200.synth import make_case, regular_grid


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        nlev, ny, nx, nens = 3, 23, 16, 5
        rng = np.random.default_rng(7)                       # same on every rank
        full_np = rng.standard_normal((nlev * ny * nx, nens))
        work = rng.uniform(1.0, 5.0, ny)
        bands = sharding.partition_bands(work, world)
        y0, y1 = bands[rank]
        dev = torch.device('cpu')

        # scatter: rank 0 supplies the state, every rank gets exactly its band
        full = torch.from_numpy(full_np.copy()) if rank == 0 else None
        mine = sharding.scatter_bands(full, bands, nlev, ny, nx, nens, torch.float64, dev, rank)
        want = full_np.reshape(nlev, ny, nx, nens)[:, y0:y1].reshape(-1, nens)
        assert np.array_equal(mine.numpy(), want), 'scatter delivered the wrong rows'

        # partial ob priors: each rank sums the stencil points it owns, all-reduce gives the full H.x
        nobs, K = 40, 8
        idx = torch.from_numpy(rng.integers(0, nlev * ny * nx, (nobs, K)))
        w = torch.from_numpy(rng.uniform(0.0, 1.0, (nobs, K)))
        lidx, lw = sharding.localize_stencil(idx, w, nlev, ny, nx, y0, y1)
        assert int(lidx.max()) < mine.shape[0] and int(lidx.min()) >= 0
        Y = (lw[:, :, None] * mine[lidx]).sum(dim=1)
        dist.all_reduce(Y)
        Y_full = (w.numpy()[:, :, None] * full_np[idx.numpy()]).sum(axis=1)
        np.testing.assert_allclose(Y.numpy(), Y_full, rtol=1e-13, atol=1e-13)

        # every rank "analyses" its band (a rank-dependent, row-dependent transform), gather rebuilds the state
        mine = mine * 2.0 + float(rank + 1)
        out = torch.zeros((nlev * ny * nx, nens), dtype=torch.float64) if rank == 0 else None
        sharding.gather_bands(mine, out, bands, nlev, ny, nx, nens, rank)
        if rank == 0:
            exp = full_np.reshape(nlev, ny, nx, nens) * 2.0
            for r, (a, b) in enumerate(bands):
                exp[:, a:b] += r + 1
            assert np.array_equal(out.numpy(), exp.reshape(-1, nens)), 'gather rebuilt the wrong state'
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:                                   # pragma: no cover
        import traceback
        q.put((rank, 'FAIL: %s\n%s' % (e, traceback.format_exc())))


@pytest.mark.timeout(180)
def test_scatter_allreduce_gather_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(msg == 'ok' for _, msg in res), res


def test_partition_bands_balances_work_and_covers_grid():
    rng = np.random.default_rng(3)
    for ny, n in ((721, 8), (181, 4), (19, 2), (8, 8)):
        work = rng.uniform(0.1, 1.0, ny)
        work[: ny // 10 + 1] *= 12.0                         # polar rows cost far more
        bands = sharding.partition_bands(work, n)
        assert bands[0][0] == 0 and bands[-1][1] == ny
        assert all(a < b for a, b in bands) and all(bands[i][1] == bands[i + 1][0] for i in range(n - 1))
        if ny >= 100:
            sums = np.array([work[a:b].sum() for a, b in bands])
            assert sums.max() <= 1.25 * sums.mean()
    with pytest.raises(ValueError):
        sharding.partition_bands(np.ones(3), 4)


def test_row_work_estimate_follows_footprints():
    """Obs uniform on the sphere give every grid point the same expected number of footprints, so rows cost
    about the same; obs clustered in one hemisphere move the work -- and the band edges -- there."""
    lat2d, lon2d = regular_grid(91, 180)
    case = make_case(ny=91, nx=180, nmem=4, nobs=400, cutoff_km=2000.0, seed=5)
    cut = 2.0 * case.ob_halfwidth
    work = sharding.estimate_row_work(lat2d, lon2d, case.ob_lat, case.ob_lon, cut, case.ob_assimilate, scan_cost=0.0)
    assert work.shape == (91,)
    assert 0.7 < work[5:15].mean() / work[40:50].mean() < 1.4
    north = case.ob_lat > 20.0
    work_n = sharding.estimate_row_work(lat2d, lon2d, case.ob_lat[north], case.ob_lon[north], cut[north],
                                        case.ob_assimilate[north], scan_cost=0.0)
    assert work_n[70:85].mean() > 10.0 * max(work_n[5:20].mean(), 1e-9)
    eq = sharding.equal_bands(91, 4)
    wb = sharding.partition_bands(work_n, 4)
    cost = lambda bands: max(work_n[a:b].sum() for a, b in bands)
    assert cost(wb) < 0.7 * cost(eq)
    assert wb[0][1] > eq[0][1]                              # the empty south gets one wide band


def test_localize_stencil_numpy_and_torch_agree():
    rng = np.random.default_rng(1)
    nlev, ny, nx = 2, 11, 7
    idx = rng.integers(0, nlev * ny * nx, (30, 8))
    w = rng.uniform(0, 1, (30, 8))
    li_n, lw_n = sharding.localize_stencil(idx, w, nlev, ny, nx, 3, 9)
    li_t, lw_t = sharding.localize_stencil(torch.from_numpy(idx), torch.from_numpy(w), nlev, ny, nx, 3, 9)
    assert np.array_equal(li_n, li_t.numpy()) and np.array_equal(lw_n, lw_t.numpy())
    # the union over a partition reproduces the global stencil exactly once
    tot = np.zeros_like(w)
    for a, b in ((0, 3), (3, 9), (9, 11)):
        tot += sharding.localize_stencil(idx, w, nlev, ny, nx, a, b)[1]
    assert np.array_equal(tot, w)


def _merge_worker(rank, world, port, q):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        from efa_xray_b200.engine import merge_distributed_records
        nobs, block = 37, 5
        rng = np.random.default_rng(3)                      # same on every rank: the complete result
        full_rec = rng.standard_normal((8, nobs))
        skipped = rng.uniform(size=nobs) < 0.2
        full_rec[2:4, skipped] = np.nan                      # post mean / variance of obs that were not assimilated
        full_ym = rng.standard_normal(nobs)
        pairs = rng.integers(1, 50, nobs)
        mine = (np.arange(nobs) // block) % world == rank
        # what a rank holds after exb_obs_solve_dist_*: its own rows, zeros elsewhere in rec, untouched input elsewhere in Ym
        rec = torch.from_numpy(np.where(mine[None, :], full_rec, 0.0))
        ym = torch.from_numpy(np.where(mine, full_ym, 123.0))
        cnt = torch.tensor([int(pairs[mine].sum())])
        merge_distributed_records(ym, rec, cnt, block, rank, world, None)
        np.testing.assert_array_equal(ym.numpy(), full_ym)
        np.testing.assert_array_equal(rec.numpy(), full_rec)          # NaNs where they were, bit-equal elsewhere
        assert int(cnt.item()) == int(pairs.sum())
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, 'FAIL: ' + traceback.format_exc()))


def test_distributed_solve_record_merge_over_gloo():
    """engine.merge_distributed_records: every rank contributes the rows dealt to it (blocks of consecutive obs), the
    sums rebuild the complete records on every rank, NaN (not assimilated) survives."""
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == 'ok' for r in results), results
