"""Drop-ins for the hot-path callables of the reference's visco/decompress_ms.py, backed by libvisco_b200.so.

  unstack_vis(vis_reconstructed, nrows)     reference decompress_ms.py:95-104
  reconstruct_vis(U, S, Vt)                 reference decompress_ms.py:107-131
  reconstruct_vis_batched(factors)          what the batch loop (decompress_ms.py:196-213) calls
"""
from __future__ import annotations

import numpy as np

from .engine import get_engine


def _np(x):
    if hasattr(x, "compute"):
        x = x.compute()
    return np.asarray(x)


def unstack_vis(vis_reconstructed, nrows):
    """Return list of blocks each with shape (nrows, nchan) (corr-optimized leaves hold vstacked correlations)."""
    vis = _np(vis_reconstructed)
    nstack = vis.shape[0] // nrows
    return list(np.split(vis, nstack, axis=0))


def _check_factors(U, S, Vt):
    U, S, Vt = _np(U), _np(S), _np(Vt)
    if S.ndim == 2:           # reference accepts (k, 1) as well as (k,)  (decompress_ms.py:125-126)
        S = S[:, 0]
    if U.ndim != 2 or Vt.ndim != 2 or S.ndim != 1 or U.shape[1] != S.shape[0] or Vt.shape[0] != S.shape[0]:
        raise ValueError(f"inconsistent factor shapes U{U.shape} S{S.shape} Vt{Vt.shape}")
    return U, S, Vt


def reconstruct_vis(U, S, Vt) -> np.ndarray:
    """(U * S[None, :]) @ Vt for one matrix: U (time, mode), S (mode,) or (mode, 1), Vt (mode, channel)."""
    U, S, Vt = _check_factors(U, S, Vt)
    if S.shape[0] == 0:
        return np.zeros((U.shape[0], Vt.shape[1]), np.complex64)
    return get_engine().reconstruct_host(U[None], S[None], Vt[None])[0]


def reconstruct_vis_batched(factors):
    """factors: list of (U, S, Vt) with common (time, channel) shape but per-matrix rank -> [B, m, n] numpy."""
    if not factors:
        return np.zeros((0, 0, 0), np.complex64)
    fs = [_check_factors(*f) for f in factors]
    m, n = fs[0][0].shape[0], fs[0][2].shape[1]
    kmax = max(1, max(f[1].shape[0] for f in fs))
    B = len(fs)
    U = np.zeros((B, m, kmax), np.complex64)
    S = np.zeros((B, kmax), np.float32)
    Vt = np.zeros((B, kmax, n), np.complex64)
    ranks = np.zeros((B,), np.int32)
    for b, (u, s, vt) in enumerate(fs):
        if u.shape[0] != m or vt.shape[1] != n:
            raise ValueError("all matrices of a batch must share (time, channel)")
        k = s.shape[0]
        U[b, :, :k], S[b, :k], Vt[b, :k, :], ranks[b] = u, s, vt, k
    return get_engine().reconstruct_host(U, S, Vt, ranks)
