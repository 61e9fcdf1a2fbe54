"""The CPU arm that bench.py times beside the GPU path (TEST/BENCH INFRASTRUCTURE, never shipped).

It restates the reference's restored pipeline (hpc/image_processing.c:183-275) with the fastest building
blocks available on the box's host cores, which is what the reference itself would get from
PETSc/SLEPc + BLAS/LAPACK (not installable here): OpenMP kernel evaluation from oracle.c (two exps per
pair, hpc/affinity.c:99,107), LAPACK dsyevd for the p x p eigensolve, BLAS dgemm for the extrapolation
(nystroem.c:41-42) and dgemv for the filter (display.c:64-73), all in fp64.

`rows=(r0, r1)` restricts the pixel-proportional stages to a band of image rows (the bounded sample of
bench.py); the p x p stages always run in full."""
from __future__ import annotations

import time

import numpy as np

from . import oracle_c as oc


def run(img, sample_indices, m=None, kind="bilateral", h_loc=40.0, h_val=30.0, gain=3.0, power=1.0, rows=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    C = 1 if img.ndim == 2 else img.shape[2]
    r0, r1 = (0, H) if rows is None else rows
    s = np.asarray(sample_indices, dtype=np.uint32)
    p = len(s)
    if m is None or m < 0 or m >= p:
        m = p - 1
    t = {}
    t_all = time.perf_counter()

    t0 = time.perf_counter()
    band = np.arange(r0 * W, r1 * W, dtype=np.uint32)
    keep = np.ones(len(band), dtype=bool)
    inside = s[(s >= r0 * W) & (s < r1 * W)]
    keep[inside - r0 * W] = False
    rest = band[keep]
    K_A = oc.affinity_rows(img, s, s, kind, h_loc, h_val)
    K_B = oc.affinity_rows(img, s, rest, kind, h_loc, h_val)          # p x nb, as the reference stores it
    t["affinity"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    D = K_A.sum(axis=1) + K_B.sum(axis=1)
    alpha = 1.0 / D.mean()
    L_A = alpha * (np.diag(D) - K_A)
    L_B = -alpha * K_B                                                  # the reference's second copy (laplacian.c:38-39)
    t["laplacian"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    mu, U = np.linalg.eigh(L_A)
    mu, U = mu[:m], U[:, :m]
    t["eigensolve"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    phi_B = L_B.T @ (U * (1.0 / mu))                                    # nystroem.c:41-42
    t["nystroem"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    y = img.reshape(H * W, C).astype(np.float64)
    c = U.T @ y[s[(s >= r0 * W) & (s < r1 * W)].astype(np.int64)] if False else None
    in_band = (s >= r0 * W) & (s < r1 * W)
    c = U[in_band].T @ y[s[in_band].astype(np.int64)] + phi_B.T @ y[rest.astype(np.int64)]
    w = (mu[:, None] ** power) * c
    z = y.copy()
    z[rest.astype(np.int64)] += gain * (phi_B @ w)
    z[s[in_band].astype(np.int64)] += gain * (U[in_band] @ w)
    z = np.minimum(z, 255.0).reshape(img.shape)
    t["filter"] = time.perf_counter() - t0
    t["total"] = time.perf_counter() - t_all
    return dict(z=z, mu=mu, D=D, alpha=alpha, timings=t, band_pixels=(r1 - r0) * W, p=p, m=m)
