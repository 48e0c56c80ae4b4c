"""Deterministic synthetic matrices of BASELINE.json's configs, host side (numpy).
Returned as COO triplets (uint64 row, uint64 col, values) in row-major emission order."""
from __future__ import annotations

import numpy as np


def laplacian_2d(g: int, dtype=np.float64):
    """C1: 5-point Laplacian on a g x g grid; row r = i*g + j; per row: diag 4, N, S, W, E = -1."""
    i, j = np.meshgrid(np.arange(g, dtype=np.int64), np.arange(g, dtype=np.int64), indexing="ij")
    r = (i * g + j).ravel()
    rows, cols, vals = [], [], []
    for di, dj, v in ((0, 0, 4.0), (-1, 0, -1.0), (1, 0, -1.0), (0, -1, -1.0), (0, 1, -1.0)):
        ii, jj = i + di, j + dj
        ok = ((ii >= 0) & (ii < g) & (jj >= 0) & (jj < g)).ravel()
        rows.append(np.where(ok, r, -1))
        cols.append(np.where(ok, (ii * g + jj).ravel(), -1))
        vals.append(np.full(g * g, v))
    # interleave so that each row's entries are emitted together, in the listed order
    R = np.stack(rows, 1).ravel()
    Cc = np.stack(cols, 1).ravel()
    V = np.stack(vals, 1).ravel()
    keep = R >= 0
    return R[keep].astype(np.uint64), Cc[keep].astype(np.uint64), V[keep].astype(dtype)


def stencil_27(m: int, dtype=np.float64):
    """C2: 27-point stencil on an m^3 grid, 26 on the diagonal, -1 elsewhere; nnz = (3m-2)^3."""
    idx = np.arange(m, dtype=np.int64)
    i, j, k = np.meshgrid(idx, idx, idx, indexing="ij")
    r = ((i * m + j) * m + k).ravel()
    rows, cols, vals = [], [], []
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for dk in (-1, 0, 1):
                ii, jj, kk = i + di, j + dj, k + dk
                ok = ((ii >= 0) & (ii < m) & (jj >= 0) & (jj < m) & (kk >= 0) & (kk < m)).ravel()
                rows.append(r[ok])
                cols.append(((ii * m + jj) * m + kk).ravel()[ok])
                vals.append(np.full(int(ok.sum()), 26.0 if (di, dj, dk) == (0, 0, 0) else -1.0))
    R, Cc, V = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((Cc, R))        # row-major, columns ascending
    return R[order].astype(np.uint64), Cc[order].astype(np.uint64), V[order].astype(dtype)


def banded(n: int, offsets, dtype=np.float64, scale=1e-3):
    """C5-style band: A[i, i+d] = 1/(1+|d|) + (i mod 7)*scale for d in offsets, clipped."""
    i = np.arange(n, dtype=np.int64)
    rows, cols, vals = [], [], []
    for d in offsets:
        ok = (i + d >= 0) & (i + d < n)
        rows.append(i[ok]); cols.append(i[ok] + d)
        vals.append(1.0 / (1 + abs(d)) + (i[ok] % 7) * scale)
    R, Cc, V = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((Cc, R))
    return R[order].astype(np.uint64), Cc[order].astype(np.uint64), V[order].astype(dtype)


def csr_from_sorted_triplets(nrows, rows, cols, vals):
    """Row-sorted, duplicate-free triplets -> (rowptr, colind, values) without any sort."""
    counts = np.bincount(rows.astype(np.int64), minlength=nrows)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    return ptr, cols.astype(np.uint64), vals


def random_coo(rng, nrows, ncols, n, dtype, dup_frac=0.05, cancel_frac=0.01):
    """Uniform random triplets with duplicates (some multiplicity >= 3) and exact cancellations."""
    r = rng.integers(0, nrows, n).astype(np.uint64)
    c = rng.integers(0, ncols, n).astype(np.uint64)
    v = rng.uniform(-1, 1, n).astype(dtype)
    nd = int(n * dup_frac)
    if nd and n:
        src = rng.integers(0, n, nd)
        r = np.concatenate([r, r[src]]); c = np.concatenate([c, c[src]])
        vv = rng.uniform(-1, 1, nd).astype(dtype)
        k = int(nd * cancel_frac / max(dup_frac, 1e-9)) if dup_frac else 0
        k = min(k, nd)
        vv[:k] = -v[src[:k]]                 # exact negations: exercises the zero drop
        v = np.concatenate([v, vv])
        # a few cells with multiplicity >= 3 (order-sensitive sums)
        m3 = max(1, nd // 10)
        src3 = src[:m3]
        r = np.concatenate([r, r[src3]]); c = np.concatenate([c, c[src3]])
        v = np.concatenate([v, (rng.uniform(-1, 1, m3) * 1e-7).astype(dtype)])
    perm = rng.permutation(len(v))
    return r[perm], c[perm], v[perm]
