"""CPU oracle: numpy restatement of VISCO's SVD -> truncate -> reconstruct path.

TEST INFRASTRUCTURE ONLY. Nothing under visco_b200/ imports this module; it is the checker the CUDA
path is compared with (tests/, __graft_entry__.smoke(), bench.py cpu_baseline and --impl reference).

What it restates (reference paths are relative to the upstream repo root):
  * visco/compress_ms.py:295-319   find_n_decorrelation   -> ref_find_n_decorrelation
  * visco/compress_ms.py:322-363   apply_svd              -> ref_apply_svd
  * visco/decompress_ms.py:95-104  unstack_vis            -> ref_unstack_vis
  * visco/decompress_ms.py:107-131 reconstruct_vis        -> ref_reconstruct_vis
The arithmetic itself lives in third-party code that is NOT under the reference tree:
  dask==2024.10.0  dask.array.linalg.svd  (single-chunk branch: np.linalg.svd(a, full_matrices=False)
                   followed by dask.array.utils.svd_flip(u, v)), da.sum / da.cumsum (np.sum / np.cumsum)
  numpy==2.2.x     np.linalg.svd -> LAPACK cgesdd (complex64)
dask is not installed in this environment, so svd_flip is restated here from its published definition.

Pinning status: the reference's own tests hold no numerical assertion for this path
(tests/compression_tests.py:35-56 and tests/decompression_tests.py:15-39 check file/column existence only).
The oracle is therefore pinned against outputs of the reference's OWN function bodies executed in this
container: tests/golden/make_golden.py extracts apply_svd / find_n_decorrelation / reconstruct_vis /
unstack_vis from the reference sources with `ast`, runs them with a numpy-backed stand-in for the missing
`dask.array` module, and commits inputs + outputs under tests/golden/. tests/test_oracle.py checks this
module against those fixtures (including matrices decoded from the reference's sample Measurement Set).
"""
from __future__ import annotations

import numpy as np


def svd_flip(u: np.ndarray, v: np.ndarray):
    """dask.array.utils.svd_flip (dask 2024.10.0), v-based decision (the default used by da.linalg.svd).

    signs[i] = +1 if sum(v[i, :]) >= 0 else -1, with numpy's lexicographic ordering for complex sums;
    column i of u and row i of v are multiplied by the same sign.
    """
    dtype = v.dtype
    signs = np.sum(v, axis=1, keepdims=True).T
    signs = 2.0 * ((signs >= 0) - 0.5).astype(dtype)
    return u * signs, v * signs.T


def ref_svd(a: np.ndarray):
    """da.linalg.svd(a) for a single-chunk 2-D array (compress_ms.py:347-350)."""
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    u, vt = svd_flip(u, vt)
    return u, s, vt


def ref_find_n_decorrelation(singular_values: np.ndarray, decorrelation: float) -> int:
    """compress_ms.py:295-319. Threshold is on the ENERGY fraction decorrelation**2; float32 arithmetic
    when the singular values are float32 (python-float * np.float32 stays float32 under NEP 50)."""
    sum_total = np.sum(singular_values ** 2)
    threshold = (decorrelation) ** 2 * sum_total
    cumulative = np.cumsum(singular_values ** 2)
    n = int(np.argmax(cumulative >= threshold)) + 1
    if n == 0:  # unreachable, kept because the reference has it (compress_ms.py:316-317)
        n = len(singular_values)
    return n


def ref_apply_svd(visdata: np.ndarray, decorrelation: float | None = None, compressionrank: int | None = None):
    """compress_ms.py:322-363: economy SVD, then fixed rank (wins) or energy rule or full rank."""
    u, s, vt = ref_svd(np.asarray(visdata))
    if compressionrank:
        n = compressionrank
    elif decorrelation:
        n = ref_find_n_decorrelation(s, decorrelation)
    else:
        n = len(s)
    return u[:, :n], s[:n], vt[:n, :]


def ref_reconstruct_vis(u: np.ndarray, s: np.ndarray, vt: np.ndarray) -> np.ndarray:
    """decompress_ms.py:107-131: (U * S[None, :]) @ Vt; S may be (k,) or (k, 1)."""
    if s.ndim == 2:
        s = s[:, 0]
    s_reshaped = s.reshape((1, s.shape[0]))
    return (u * s_reshaped) @ vt


def ref_unstack_vis(vis_reconstructed: np.ndarray, nrows: int):
    """decompress_ms.py:95-104: split a vstacked (corr-optimized) reconstruction into (nrows, nchan) blocks."""
    nstack = vis_reconstructed.shape[0] // nrows
    return list(np.split(vis_reconstructed, nstack, axis=0))


# ----------------------------------------------------------------------------------------------------
# helpers used by tests / bench (not part of the reference)
# ----------------------------------------------------------------------------------------------------
def roundtrip(a: np.ndarray, decorrelation=None, compressionrank=None):
    """compress + reconstruct one matrix the way the reference does; returns (recon, s, k)."""
    u, s, vt = ref_apply_svd(a, decorrelation, compressionrank)
    return ref_reconstruct_vis(u, s, vt), s, len(s)
