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


def leaf_planes(zarr_path: str, ncorr: int) -> dict:
    """Leaf directory name -> correlation plane(s) of the (row, chan, corr) column.

    The compressor picks planes by looking the casacore enum up in POLARIZATION/CORR_TYPE (reference compress_ms.py:
    601-602, 631-632, 662), so the decompressor does the same when the store carries that table: a 2-correlation
    [XX, YY] column then gets its "YY" / "diagonals" leaves back in planes (1) / (0, 1). The reference's decompressor
    hard-codes linear-feed positions instead (decompress_ms.py:182, 222-229: XX 0, XY 1, YX 2, YY -1, diagonals 0/3,
    offdiagonals 1/2) - identical for the usual [9, 10, 11, 12] column, and used here when CORR_TYPE is absent."""
    import os

    from .msdata import CORR_TYPES
    from .zarr_leaf import read_array
    planes = {"XX": (0,), "XY": (1,), "YX": (2,), "YY": (ncorr - 1,), "diagonals": (0, 3), "offdiagonals": (1, 2)}
    p = os.path.join(zarr_path, "POLARIZATION", "CORR_TYPE")
    if os.path.isdir(p):
        ct = [int(x) for x in np.asarray(read_array(p)).reshape(-1)[:ncorr]]
        by_name = {name: (ct.index(enum),) for name, enum in CORR_TYPES.items() if enum in ct}
        planes.update(by_name)
        if 9 in ct and 12 in ct:
            planes["diagonals"] = (ct.index(9), ct.index(12))
        if 10 in ct and 11 in ct:
            planes["offdiagonals"] = (ct.index(10), ct.index(11))
    return planes


def construct_main_ds(zarr_path: str, column: str, batch_size: int, ngpus: int = 1):
    """Rebuild the visibility column from the leaf tree (reference construct_main_ds, decompress_ms.py:134-234).

    Device pipeline: per batch the zero-padded factors go to the GPU, vk_reconstruct_batched forms the matrices and
    vk_scatter_baselines writes them straight into the (row, chan, corr) column held on the device (including the
    unstacking of "diagonals" / "offdiagonals" leaves); the column comes back to the host once at the end.
    ngpus > 1 (additive keyword): the reconstruction tasks are split into one contiguous range per GPU, one host thread
    and handle each; every GPU scatters into its own zero-initialised copy of the column and the copies - disjoint by
    construction - are summed on the host.
    Returns a visco_b200.msdata.VisData whose `data` is the reconstructed [row, chan, corr] complex64 array."""
    import os
    import threading

    import torch

    from . import LOG
    from .engine import Engine
    from .msdata import VisData
    from .shard import shard_baselines
    from .zarr_leaf import list_subtables, read_svd_from_zarr
    eng = get_engine()
    dev = f"cuda:{eng.device}"
    ant1, ant2, rowid, antnames, shape = _store_index(zarr_path)
    base = os.path.join(zarr_path, "MAIN", column)
    tasks = []                                         # (U, S, WT, row_indices, corr planes, stack)
    nchan = None
    for baseline in list_subtables(base):
        correlations = list_subtables(os.path.join(base, baseline))
        if "&" not in baseline or not correlations:
            continue
        a1n, a2n = baseline.split("&")
        if a1n not in antnames or a2n not in antnames:
            LOG.warning(f"Baseline {baseline} not found in ANTENNA table. Skipping.")      # reference :175-177
            continue
        row_indices = np.nonzero((ant1 == antnames.index(a1n)) & (ant2 == antnames.index(a2n)))[0]
        for corr_name in correlations:
            U, S, WT, _ = read_svd_from_zarr(os.path.join(base, baseline, corr_name))
            nchan = WT.shape[1]
            tasks.append((U, S, WT, row_indices, corr_name))
    if shape is None:
        if nchan is None:
            raise ValueError(f"{zarr_path} holds no factor leaves under MAIN/{column}")
        shape = [len(ant1), nchan, 4]
    ncorr = int(shape[2])
    planes = leaf_planes(zarr_path, ncorr)
    batch_size = max(1, int(batch_size))
    for t in tasks:
        if t[4] not in planes:
            raise ValueError(f"unknown leaf name {t[4]}")
        if any(p < 0 or p >= ncorr for p in planes[t[4]]):
            # the reference's numpy assignment raises IndexError here (decompress_ms.py:222-229 on a 2-correlation column)
            raise ValueError(f"leaf {t[4]} maps to correlation plane(s) {planes[t[4]]} but the column has {ncorr}")
        if t[0].shape[0] != len(planes[t[4]]) * t[3].size:
            raise ValueError(f"leaf {t[4]} has {t[0].shape[0]} rows, the table has {t[3].size} for this baseline")

    def run(eng_g, my_tasks):
        dev_g = f"cuda:{eng_g.device}"
        with torch.cuda.device(eng_g.device):
            out_g = torch.zeros(tuple(int(x) for x in shape), dtype=torch.complex64, device=dev_g)
            for start in range(0, len(my_tasks), batch_size):
                groups = {}
                for t in my_tasks[start:start + batch_size]:
                    groups.setdefault((t[3].size, t[2].shape[1], len(planes[t[4]])), []).append(t)
                for (m, n, stack), ts in groups.items():
                    B = len(ts)
                    kmax = max(1, max(t[1].shape[0] for t in ts))
                    U = np.zeros((B, stack * m, kmax), np.complex64)
                    S = np.zeros((B, kmax), np.float32)
                    Vt = np.zeros((B, kmax, n), np.complex64)
                    for b, t in enumerate(ts):
                        k = t[1].shape[0]
                        U[b, :, :k], S[b, :k], Vt[b, :k, :] = t[0], t[1], t[2]
                    rec = eng_g.reconstruct(torch.from_numpy(U).to(dev_g), torch.from_numpy(S).to(dev_g), torch.from_numpy(Vt).to(dev_g))
                    row_idx = torch.from_numpy(np.stack([t[3] for t in ts]).astype(np.int32)).to(dev_g)
                    corr_sel = torch.tensor([list(planes[t[4]]) for t in ts], dtype=torch.int32, device=dev_g)
                    eng_g.scatter_baselines(rec, out_g, row_idx, corr_sel, stack)
            return out_g

    ngpus = max(1, min(int(ngpus or 1), torch.cuda.device_count(), max(1, len(tasks))))
    if ngpus == 1:
        out_dev = run(eng, tasks)
    else:
        parts, errors = [None] * ngpus, []

        def worker(g):
            try:
                off, cnt = shard_baselines(len(tasks), ngpus, g)
                e_g = eng if g == eng.device else Engine(g)
                parts[g] = run(e_g, tasks[off:off + cnt]).cpu()
            except BaseException as ex:
                errors.append(ex)
        threads = [threading.Thread(target=worker, args=(g,)) for g in range(ngpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        out_dev = parts[0]
        for p_ in parts[1:]:
            out_dev += p_
    out = out_dev.cpu().numpy()
    # flags: unpack the bit-packed FLAGS / FLAGS_ROW groups (reference :240-246)
    flag = flag_row = None
    from .zarr_leaf import read_array
    if os.path.isdir(os.path.join(zarr_path, "FLAGS", "FLAGS")):
        packed = torch.from_numpy(read_array(os.path.join(zarr_path, "FLAGS", "FLAGS")).astype(np.uint8)).to(dev)
        flag = eng.unpackbits(packed, out.size).cpu().numpy().astype(bool).reshape(out.shape)
    if os.path.isdir(os.path.join(zarr_path, "FLAGS_ROW", "FLAGS_ROW")):
        packed = torch.from_numpy(read_array(os.path.join(zarr_path, "FLAGS_ROW", "FLAGS_ROW")).astype(np.uint8)).to(dev)
        flag_row = eng.unpackbits(packed, out.shape[0]).cpu().numpy().astype(bool)
    corr_types = [9, 10, 11, 12][:ncorr] if ncorr <= 4 else list(range(ncorr))
    if os.path.isdir(os.path.join(zarr_path, "POLARIZATION", "CORR_TYPE")):
        ct = [int(x) for x in np.asarray(read_array(os.path.join(zarr_path, "POLARIZATION", "CORR_TYPE"))).reshape(-1)]
        if len(ct) >= ncorr:
            corr_types = ct[:ncorr]
    # WEIGHT_SPECTRUM / SIGMA_SPECTRUM: the reference multiplies U by diag(S) ONLY (decompress_ms.py:252-254: no WT),
    # expands a trailing axis and tiles it over the correlations, i.e. the columns come back as (row, 1, corr) with the
    # row profile of the weights and no channel dependence; both columns get the same array (:257-270). Kept as is.
    weights = None
    if os.path.isdir(os.path.join(zarr_path, "WEIGHT_SPECTRUM", "U")):
        Uw, Sw, _, _ = read_svd_from_zarr(os.path.join(zarr_path, "WEIGHT_SPECTRUM"))
        wrec = np.dot(Uw.real.astype(np.float32), np.diag(Sw))
        weights = np.tile(np.expand_dims(wrec, axis=-1), (1, 1, ncorr))
    return VisData(data=out, antenna1=ant1, antenna2=ant2, antenna_names=antnames, corr_types=corr_types, rowid=rowid,
                   flag=flag, flag_row=flag_row, weight_spectrum=weights, sigma_spectrum=weights)


def open_dataset(zarr_path: str, column: str = "COMPRESSED_DATA", batch_size: int = 50, ngpus: int = 1):
    """Decompress to an in-memory data set without writing an MS (reference open_dataset, decompress_ms.py:295-326)."""
    return construct_main_ds(zarr_path, column, batch_size, ngpus=ngpus)


def write_datasets_to_ms(zarr_path: str, msname: str, column: str, batch_size: int, ngpus: int = 1):
    """Decompress a store into `msname` (reference write_datasets_to_ms, decompress_ms.py:329-402). Writing a casacore
    Measurement Set needs dask-ms / python-casacore (out of scope, absent here): an ``.npz`` bundle name is written
    directly, anything else raises unless python-casacore is importable."""
    vis = construct_main_ds(zarr_path, column, batch_size, ngpus=ngpus)
    if str(msname).endswith(".npz"):
        vis.save(msname)
        return msname
    try:
        from casacore.tables import table  # type: ignore  # noqa: F401
    except ImportError as e:
        raise RuntimeError("writing a Measurement Set needs python-casacore, which is not installed; "
                           "use an .npz output name") from e
    raise NotImplementedError("MS writing through python-casacore is outside the scope of this build (SURVEY 2.1 row 7)")
