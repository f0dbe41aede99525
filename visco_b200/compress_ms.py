"""Drop-ins for the hot-path callables of the reference's visco/compress_ms.py, backed by libvisco_b200.so.

  find_n_decorrelation(singular_values, decorrelation)      reference compress_ms.py:295-319
  apply_svd(visdata, decorrelation=None, compressionrank=None)   reference compress_ms.py:322-363
  apply_svd_batched(cube, ...)   what the batch loop (compress_ms.py:571-697) calls instead of one task per matrix

Differences a maintainer should know about (all on purpose):
  * results are numpy arrays, not lazy dask arrays (the reference re-evaluates the lazy SVD 3-5x per matrix);
    `.compute()` is not needed by write_svd_to_zarr any more.
  * complex singular vectors are unique only up to a phase per (u_i, v_i) pair; ours differ from LAPACK's by such
    phases. S, the reconstruction U*S@Vt and the chosen rank agree with the reference within the tolerances of
    BASELINE.json (1e-4 on S, 1e-5 on the reconstruction error, ranks equal away from threshold ties).
  * input is computed in complex64 (the dtype of MS DATA columns); other dtypes are cast.
There is no CPU fallback: without a B200 these functions raise.
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine


def _as_matrix(visdata) -> np.ndarray:
    if hasattr(visdata, "compute"):  # a dask array handed in by unchanged reference code
        visdata = visdata.compute()
    a = np.asarray(visdata)
    if a.ndim != 2:
        raise ValueError(f"apply_svd expects a 2-D (time x channel) matrix, got shape {a.shape}")
    return np.ascontiguousarray(a, dtype=np.complex64)


def find_n_decorrelation(singular_values, decorrelation: float) -> int:
    """Number of singular values whose cumulative energy reaches decorrelation**2 of the total (float32
    arithmetic, as the reference computes it). Runs on the GPU like every other hot-path function."""
    import torch
    if hasattr(singular_values, "compute"):
        singular_values = singular_values.compute()
    s = np.ascontiguousarray(np.asarray(singular_values), dtype=np.float32).reshape(1, -1)
    eng = get_engine()
    ranks = eng.find_n_decorrelation(torch.from_numpy(s).to(f"cuda:{eng.device}"), float(decorrelation))
    return int(ranks.cpu()[0])


def apply_svd(visdata, decorrelation: float = None, compressionrank: int = None):
    """Decompose one visibility matrix; returns (U[:, :n], S[:n], Vt[:n, :]) as numpy arrays."""
    a = _as_matrix(visdata)
    U, S, Vt, ranks, _ = get_engine().compress_host(a[None], decorrelation, compressionrank)
    k = int(ranks[0])
    return U[0, :, :k].copy(), S[0, :k].copy(), Vt[0, :k, :].copy()


def apply_svd_batched(cube, decorrelation: float = None, compressionrank: int = None):
    """cube: [B, m, n] complex64 (numpy). Returns a list of B (U, S, Vt) tuples truncated to each matrix's rank."""
    cube = np.ascontiguousarray(np.asarray(cube), dtype=np.complex64)
    if cube.ndim != 3:
        raise ValueError(f"apply_svd_batched expects [B, time, channel], got shape {cube.shape}")
    if cube.shape[0] == 0:
        return []
    U, S, Vt, ranks, _ = get_engine().compress_host(cube, decorrelation, compressionrank)
    return [(U[b, :, :k].copy(), S[b, :k].copy(), Vt[b, :k, :].copy()) for b, k in enumerate(ranks.tolist())]
