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


# =====================================================================================================================
# Callers on the decompression side of the hot path: leaf discovery, batched reconstruction, scatter into
# (row, chan, corr). Mirrors reference decompress_ms.py:134-402 with one reconstruct_vis_batched call per batch.
# =====================================================================================================================
def _store_index(zarr_path):
    """ANTENNA1, ANTENNA2, ROWID, antenna names and the shape of the visibility column of a compressed store.
    Works on stores written by visco_b200 and on stores written by the reference (zarr-v2 groups MAIN / ANTENNA whose
    DATA chunks were deleted but whose metadata survives, reference compress_ms.py:934-939, decompress_ms.py:151-161)."""
    import json
    import os

    from .zarr_leaf import read_array
    main = os.path.join(zarr_path, "MAIN")
    ant1 = read_array(os.path.join(main, "ANTENNA1"))
    ant2 = read_array(os.path.join(main, "ANTENNA2"))
    rowid = read_array(os.path.join(main, "ROWID")) if os.path.isdir(os.path.join(main, "ROWID")) else np.arange(len(ant1))
    names = [str(x) for x in read_array(os.path.join(zarr_path, "ANTENNA", "NAME"))]
    shape = None
    attrs_p = os.path.join(main, ".zattrs")
    if os.path.exists(attrs_p):
        shape = json.load(open(attrs_p)).get("visco_b200", {}).get("data_shape")
    if shape is None:
        for meta_p, key in ((os.path.join(zarr_path, ".zmetadata"), "MAIN/DATA/.zarray"),
                            (os.path.join(main, ".zmetadata"), "DATA/.zarray")):
            if os.path.exists(meta_p):
                md = json.load(open(meta_p)).get("metadata", {})
                if key in md:
                    shape = md[key]["shape"]
                    break
    if shape is None and os.path.exists(os.path.join(main, "DATA", ".zarray")):
        shape = json.load(open(os.path.join(main, "DATA", ".zarray")))["shape"]
    return ant1, ant2, rowid, names, shape


def construct_main_ds(zarr_path: str, column: str, batch_size: int):
    """Rebuild the visibility column from the leaf tree (reference construct_main_ds, decompress_ms.py:134-234).
    Returns a visco_b200.msdata.VisData whose `data` is the reconstructed [row, chan, corr] complex64 array."""
    import os

    from . import LOG
    from .msdata import VisData
    from .zarr_leaf import list_subtables, read_svd_from_zarr
    ant1, ant2, rowid, antnames, shape = _store_index(zarr_path)
    base = os.path.join(zarr_path, "MAIN", column)
    baselines = list_subtables(base)
    tasks = []                                         # (U, S, WT, row_indices, corr_name)
    nchan = None
    for baseline in baselines:
        correlations = list_subtables(os.path.join(base, baseline))
        if "&" not in baseline or not correlations:
            continue
        a1n, a2n = baseline.split("&")
        if a1n not in antnames or a2n not in antnames:
            LOG.warning(f"Baseline {baseline} not found in ANTENNA table. Skipping.")      # reference :175-177
            continue
        a1, a2 = antnames.index(a1n), antnames.index(a2n)
        row_indices = np.nonzero((ant1 == a1) & (ant2 == a2))[0]
        for corr_name in correlations:
            U, S, WT, _ = read_svd_from_zarr(os.path.join(base, baseline, corr_name))
            nchan = WT.shape[1]
            tasks.append((U, S, WT, row_indices, corr_name))
    if shape is None:
        if nchan is None:
            raise ValueError(f"{zarr_path} holds no factor leaves under MAIN/{column}")
        shape = [len(ant1), nchan, 4]
    out = np.zeros(tuple(shape), dtype=np.complex64)
    corr_indices = {"XX": 0, "XY": 1, "YX": 2, "YY": -1}                                  # reference :182
    batch_size = max(1, int(batch_size))
    for start in range(0, len(tasks), batch_size):
        batch = tasks[start:start + batch_size]
        by_shape = {}
        for j, t in enumerate(batch):
            by_shape.setdefault((t[0].shape[0], t[2].shape[1]), []).append(j)
        for _, idx in by_shape.items():
            rec = reconstruct_vis_batched([batch[j][:3] for j in idx])
            for j, vis in zip(idx, rec):
                _, _, _, row_indices, corr_name = batch[j]
                nrows = row_indices.size
                if corr_name == "diagonals":                                             # reference :222-225
                    parts = unstack_vis(vis, nrows)
                    out[row_indices, :, 0] = parts[0]
                    out[row_indices, :, 3] = parts[1]
                elif corr_name == "offdiagonals":                                        # reference :226-229
                    parts = unstack_vis(vis, nrows)
                    out[row_indices, :, 1] = parts[0]
                    out[row_indices, :, 2] = parts[1]
                else:
                    out[row_indices, :, corr_indices[corr_name]] = vis
    corr_types = [9, 10, 11, 12][: out.shape[2]] if out.shape[2] <= 4 else list(range(out.shape[2]))
    return VisData(data=out, antenna1=ant1, antenna2=ant2, antenna_names=antnames, corr_types=corr_types, rowid=rowid)


def open_dataset(zarr_path: str, column: str = "COMPRESSED_DATA", batch_size: int = 50):
    """Decompress to an in-memory data set without writing an MS (reference open_dataset, decompress_ms.py:295-326)."""
    return construct_main_ds(zarr_path, column, batch_size)


def write_datasets_to_ms(zarr_path: str, msname: str, column: str, batch_size: int):
    """Decompress a store into `msname` (reference write_datasets_to_ms, decompress_ms.py:329-402). Writing a casacore
    Measurement Set needs dask-ms / python-casacore (out of scope, absent here): an ``.npz`` bundle name is written
    directly, anything else raises unless python-casacore is importable."""
    vis = construct_main_ds(zarr_path, column, batch_size)
    if str(msname).endswith(".npz"):
        vis.save(msname)
        return msname
    try:
        from casacore.tables import table  # type: ignore  # noqa: F401
    except ImportError as e:
        raise RuntimeError("writing a Measurement Set needs python-casacore, which is not installed; "
                           "use an .npz output name") from e
    raise NotImplementedError("MS writing through python-casacore is outside the scope of this build (SURVEY 2.1 row 7)")
