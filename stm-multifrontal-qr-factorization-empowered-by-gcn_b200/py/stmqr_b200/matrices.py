"""Synthetic workload generators named by BASELINE.json / SURVEY.md 8(d).

All return (m, n, Ap, Ai, Ax) in CSC with int64 indices and sorted rows.
"""
from __future__ import annotations

import numpy as np


def _coo_to_csc(m, n, rows, cols, vals):
    import scipy.sparse as sp
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return m, n, A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)


def laplacian_2d(nx: int, ny: int | None = None):
    """5-point Laplacian, natural numbering u = x + nx*y, diag 4, off-diag -1, Dirichlet
    (nx = ny = 1024)."""
    ny = nx if ny is None else ny
    idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [np.full(nx * ny, 4.0)]
    for a, b in ((idx[:, :-1], idx[:, 1:]), (idx[:-1, :], idx[1:, :])):
        rows += [a.ravel(), b.ravel()]
        cols += [b.ravel(), a.ravel()]
        vals += [np.full(a.size, -1.0), np.full(a.size, -1.0)]
    return _coo_to_csc(nx * ny, nx * ny, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def laplacian_3d(nx: int, ny: int | None = None, nz: int | None = None):
    """7-point Laplacian, u = x + nx*(y + ny*z), diag 6, off-diag -1 (config 5: 96^3)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [np.full(idx.size, 6.0)]
    for a, b in ((idx[:, :, :-1], idx[:, :, 1:]), (idx[:, :-1, :], idx[:, 1:, :]), (idx[:-1], idx[1:])):
        rows += [a.ravel(), b.ravel()]
        cols += [b.ravel(), a.ravel()]
        vals += [np.full(a.size, -1.0), np.full(a.size, -1.0)]
    N = idx.size
    return _coo_to_csc(N, N, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def tall_banded_random(m: int, n: int, draws: int = 8, halfwidth: int = 64, seed: int = 4):
    """Config 4 restated: row i gets `draws` columns (floor(i*n/m) + d) mod n,
    d ~ U{-halfwidth..halfwidth}, duplicates merged, values ~ N(0,1), PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = (np.arange(m, dtype=np.int64) * n) // m
    d = rng.integers(-halfwidth, halfwidth + 1, size=(m, draws), dtype=np.int64)
    cols = (base[:, None] + d) % n
    rows = np.repeat(np.arange(m, dtype=np.int64), draws)
    vals = rng.standard_normal(m * draws)
    # merge duplicates by keeping the first draw (sum_duplicates would change the distribution)
    key = rows * n + cols.ravel()
    _, first = np.unique(key, return_index=True)
    return _coo_to_csc(m, n, rows[first], cols.ravel()[first], vals[first])


def random_sparse(m: int, n: int, density: float, seed: int = 0, rank_deficient_cols: int = 0):
    """Small random test matrix; optionally duplicates some columns to force dead pivots."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=density, random_state=rng, format="csc", dtype=np.float64)
    A = A + sp.eye(m, n, format="csc") * 1.0 if m >= n else A
    A = A.tolil()
    for k in range(rank_deficient_cols):
        src = int(rng.integers(0, n))
        dst = int(rng.integers(0, n))
        if src != dst:
            A[:, dst] = A[:, src]
    A = A.tocsc()
    A.sort_indices()
    return m, n, A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)


def dense_random(m: int, n: int, seed: int = 7):
    """A dense m-by-n matrix stored as sparse_csc: the symbolic analysis yields ONE front of
    m x n, so the numeric phase is exactly the dense staircase-free front QR (the large-front
    kernels measured in isolation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.arange(0, (n + 1) * m, m, dtype=np.int64)
    i = np.tile(np.arange(m, dtype=np.int64), n)
    return m, n, p, i, rng.standard_normal(m * n)
