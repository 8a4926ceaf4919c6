"""Latitude-band sharding of the state across the GPUs of one box.

State rows never read other state rows, and the obs-space rows are a closed subsystem that every rank
replicates, so the per-observation loop needs no communication.  Collectives (torch.distributed, NCCL on
GPUs / gloo in the CPU tests) appear only in three places:
    scatter_bands   rank 0's full state  -> one latitude band per rank      (point-to-point sends)
    all_reduce      partial ob priors H.x -> identical full ob priors on every rank
    gather_bands    analysis bands       -> rank 0                           (point-to-point sends)

Bands are balanced by WORK, not by rows: polar rows intersect far more localisation footprints than
equatorial ones on a regular lat-lon grid.
"""
from __future__ import annotations

import numpy as np


def estimate_row_work(lat2d, lon2d, ob_lat, ob_lon, ob_cutoff_km, ob_assimilate=None, max_obs=4096,
                      col_stride=4, scan_cost=0.002):
    """Relative cost of each grid row y: (number of (ob, grid point) pairs inside the localisation cutoff,
    estimated from a sample of obs and every col_stride-th column) + a constant per grid point for the
    candidate scan.  ob_cutoff_km is the support radius (2 x half-width); inf = no localisation."""
    R = 6371.0
    ny, nx = lat2d.shape
    sel = np.arange(len(ob_lat)) if ob_assimilate is None else np.flatnonzero(ob_assimilate)
    if sel.size == 0:
        return np.ones(ny)
    if sel.size > max_obs:
        sel = sel[np.linspace(0, sel.size - 1, max_obs).astype(np.int64)]
    scale = (np.count_nonzero(ob_assimilate) if ob_assimilate is not None else len(ob_lat)) / sel.size
    cols = np.arange(0, nx, col_stride)
    phi = np.radians(lat2d[:, cols])
    lam = np.radians(lon2d[:, cols])
    gu = np.stack([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)], axis=-1)   # [ny, nc, 3]
    op, ol = np.radians(ob_lat[sel]), np.radians(ob_lon[sel])
    ou = np.stack([np.cos(op) * np.cos(ol), np.cos(op) * np.sin(ol), np.sin(op)], axis=0)          # [3, ns]
    theta = np.minimum(np.asarray(ob_cutoff_km, dtype=np.float64)[sel] / R, np.pi)
    cost = np.cos(theta)
    work = np.zeros(ny)
    for y0 in range(0, ny, 32):
        d = gu[y0:y0 + 32].reshape(-1, 3).astype(np.float32) @ ou.astype(np.float32)               # [rows*nc, ns]
        inside = (d >= cost[None, :].astype(np.float32)).reshape(-1, len(cols), sel.size)
        work[y0:y0 + 32] = inside.sum(axis=(1, 2))
    work = work * (nx / len(cols)) * scale
    nassim = scale * sel.size
    return work + scan_cost * nassim * nx


def partition_bands(work, nranks):
    """Contiguous row ranges [(y0, y1)] with near-equal summed work; every rank gets at least one row."""
    ny = len(work)
    if nranks > ny:
        raise ValueError('more ranks (%d) than grid rows (%d)' % (nranks, ny))
    csum = np.concatenate([[0.0], np.cumsum(np.asarray(work, dtype=np.float64))])
    total = csum[-1]
    edges = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        y = int(np.searchsorted(csum, target, side='left'))
        if y > 0 and abs(csum[y - 1] - target) < abs(csum[y] - target):
            y -= 1
        y = max(y, edges[-1] + 1)
        y = min(y, ny - (nranks - r))
        edges.append(y)
    edges.append(ny)
    return [(edges[i], edges[i + 1]) for i in range(nranks)]


def equal_bands(ny, nranks):
    edges = np.linspace(0, ny, nranks + 1).round().astype(int)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(nranks)]


def band_view(X, nlev, ny, nx, y0, y1):
    """View of rows y0..y1 of a full state tensor/array X[nlev*ny*nx, nens] as [nlev, y1-y0, nx, nens]."""
    return X.reshape(nlev, ny, nx, X.shape[-1])[:, y0:y1]


def localize_stencil(idx, w, nlev, ny, nx, y0, y1):
    """Re-index a global stencil (row indices into the full state) for the band [y0, y1): rows outside the
    band get weight 0 (and index 0).  Works on torch tensors or numpy arrays."""
    npts = ny * nx
    lev = idx // npts
    rem = idx - lev * npts
    y = rem // nx
    x = rem - y * nx
    inside = (y >= y0) & (y < y1)
    local = (lev * (y1 - y0) + (y - y0)) * nx + x
    local = local * inside
    return local, w * inside


def scatter_bands(X_full, bands, nlev, ny, nx, nens, dtype, device, rank, src=0, group=None):
    """Every rank returns its band [nlev*(y1-y0)*nx, nens]; rank `src` supplies X_full."""
    import torch
    import torch.distributed as dist
    y0, y1 = bands[rank]
    mine = torch.empty((nlev * (y1 - y0) * nx, nens), dtype=dtype, device=device)
    if rank == src:
        reqs, keep = [], []
        for r, (a, b) in enumerate(bands):
            part = band_view(X_full, nlev, ny, nx, a, b).contiguous().reshape(-1, nens)
            if r == src:
                mine.copy_(part)
            else:
                keep.append(part)
                reqs.append(dist.isend(part, dst=r, group=group))
        for q in reqs:
            q.wait()
    else:
        dist.recv(mine, src=src, group=group)
    return mine


def gather_bands(band, X_full, bands, nlev, ny, nx, nens, rank, dst=0, group=None):
    """Inverse of scatter_bands: rank `dst` writes every band into X_full (in place)."""
    import torch
    import torch.distributed as dist
    if rank == dst:
        full = X_full.reshape(nlev, ny, nx, nens)
        for r, (a, b) in enumerate(bands):
            if r == dst:
                full[:, a:b].copy_(band.reshape(nlev, b - a, nx, nens))
            else:
                buf = torch.empty((nlev, b - a, nx, nens), dtype=band.dtype, device=band.device)
                dist.recv(buf, src=r, group=group)
                full[:, a:b].copy_(buf)
    else:
        dist.send(band.contiguous(), dst=dst, group=group)
