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

from . import LOG
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


# =====================================================================================================================
# Callers on the compression side of the hot path: baseline discovery, batching,
# per-(baseline, correlation) gather, leaf tree. Mirrors reference compress_ms.py:366-703 with the per-matrix dask
# tasks replaced by ONE apply_svd_batched call per batch of baselines.
# =====================================================================================================================
def weight_spectrum_rank1(weight_spectrum):
    """Leading singular triplet of WEIGHT_SPECTRUM[:, :, 0] (reference compress_ms.py:493-494:
    ``apply_svd(ws, compressionrank=1)`` on the real nrow x nchan plane). The device path works in complex64; for a real
    matrix the leading pair is real up to one unit phase, which is removed here, and the sign follows dask's svd_flip
    (sum of the right vector >= 0) as the reference's factors do. Returns float32 (U[nrow, 1], S[1], Vt[1, nchan])."""
    ws = np.asarray(weight_spectrum)
    if ws.ndim == 3:
        ws = ws[:, :, 0]
    U, S, Vt = apply_svd(np.ascontiguousarray(ws, dtype=np.complex64), compressionrank=1)
    i = int(np.argmax(np.abs(U[:, 0])))
    ph = U[i, 0] / abs(U[i, 0]) if abs(U[i, 0]) > 0 else 1.0
    u = (U[:, 0] * np.conj(ph)).real.astype(np.float32)
    v = (Vt[0] * ph).real.astype(np.float32)
    if v.sum() < 0:
        u, v = -u, -v
    return u[:, None], S.astype(np.float32), v[None, :]


def batch_baselines(baselines, batch_size):
    """Split the baseline list into batches of `batch_size` (reference compress_ms.py:366-386)."""
    batch_size = max(1, int(batch_size))
    return [baselines[i:i + batch_size] for i in range(0, len(baselines), batch_size)]


def _leaf_jobs(vis, baseline_batch, correlation, correlation_optimized):
    """Yield (leaf name parts, matrix, rowid) for every SVD of a batch (reference compress_ms.py:588-688)."""
    corr_names = [c.strip() for c in correlation.split(",") if c.strip()]
    for a1, a2 in baseline_batch:
        rows = vis.baseline_rows(int(a1), int(a2))
        if rows.size == 0:
            continue
        rowid = vis.rowid[rows]
        name = f"{vis.antenna_names[int(a1)]}&{vis.antenna_names[int(a2)]}"
        block = vis.data[rows]                                  # [time, chan, corr] of this baseline
        if correlation_optimized:
            # XX and YY stacked -> "diagonals", XY and YX stacked -> "offdiagonals" (reference :600-657; the reference
            # hard-codes the casacore enums 9/12 and 10/11 here)
            if "XX" in corr_names and "YY" in corr_names:
                m = np.vstack([block[:, :, vis.corr_index(9)], block[:, :, vis.corr_index(12)]])
                yield (name, "diagonals"), m, np.tile(rowid, 2)
            if "XY" in corr_names and "YX" in corr_names:
                m = np.vstack([block[:, :, vis.corr_index(10)], block[:, :, vis.corr_index(11)]])
                yield (name, "offdiagonals"), m, np.tile(rowid, 2)
        else:
            for c in corr_names:
                yield (name, c), block[:, :, vis.corr_index(c)], rowid


def compress_visdata(vis, zarr_output_path, correlation="XX,YY", correlation_optimized=False, decorrelation=None,
                     compressionrank=None, outcolumn="COMPRESSED_DATA", compressor="zstd", level=4, batch_size=20,
                     antennas=None, data_dev=None, ngpus=1):
    """Compress every (baseline, correlation) matrix of `vis` (a visco_b200.msdata.VisData) and write the leaf tree
    ``<zarr>/MAIN/<outcolumn>/<ANT1>&<ANT2>/<corr>/`` (reference compress_visdata, compress_ms.py:389-703).

    Device pipeline: the visibility column is uploaded once; per batch of baselines the row indices go to the GPU,
    vk_gather_baselines builds the [B, time, chan] cube (all selected correlations in one pass over the rows,
    including the --correlation-optimized vstack), vk_compress_batched factorises it, and only the truncated factors
    come back to the host to be written as leaves.

    ngpus > 1: the baselines are split with shard.shard_baselines into one contiguous range per GPU; one host thread per
    GPU (its own vk_handle) runs the same batch loop on its range and writes its leaves concurrently - the role of the
    reference's ``nworkers`` dask processes (compress_ms.py:571-697, visco/__init__.py:35-89). No data-path collective.
    Returns the number of baselines processed."""
    import threading
    from pathlib import Path

    import torch

    from .engine import Engine
    from .shard import shard_baselines
    from .zarr_leaf import write_svd_to_zarr
    corr_names = [c.strip() for c in correlation.split(",") if c.strip()]
    if correlation_optimized:
        # XX+YY -> "diagonals", XY+YX -> "offdiagonals"; the reference hard-codes enums 9/12 and 10/11 (:600-657)
        leaf_names, sel = [], []
        if "XX" in corr_names and "YY" in corr_names:
            leaf_names.append("diagonals")
            sel += [vis.corr_index(9), vis.corr_index(12)]
        if "XY" in corr_names and "YX" in corr_names:
            leaf_names.append("offdiagonals")
            sel += [vis.corr_index(10), vis.corr_index(11)]
        stack = 2
    else:
        leaf_names, sel, stack = corr_names, [vis.corr_index(c) for c in corr_names], 1
    baselines = vis.baselines(antennas)
    if not leaf_names:
        return 0

    def run(eng, my_baselines, data_dev):
        dev = f"cuda:{eng.device}"
        with torch.cuda.device(eng.device):
            if data_dev is None:
                data_dev = torch.from_numpy(vis.data).to(dev)
            done = 0
            for batch in batch_baselines(my_baselines, batch_size):
                by_m = {}
                for a1, a2 in batch:
                    rows = vis.baseline_rows(int(a1), int(a2))
                    if rows.size:
                        by_m.setdefault(rows.size, []).append((int(a1), int(a2), rows))
                for m, ents in by_m.items():           # baselines of a batch may have different numbers of rows
                    row_idx = torch.from_numpy(np.stack([e[2] for e in ents]).astype(np.int32)).to(dev)
                    corr_sel = torch.tensor([sel] * len(ents), dtype=torch.int32, device=dev)
                    cube = eng.gather_baselines(data_dev, row_idx, corr_sel, stack)
                    U, S, Vt, ranks, stats = eng.compress(cube, decorrelation, compressionrank)
                    Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (U, S, Vt, ranks, stats))
                    if not np.all(st[:, 3] == 1):
                        raise np.linalg.LinAlgError("SVD did not converge")
                    for e, (a1, a2, rows) in enumerate(ents):
                        name = f"{vis.antenna_names[a1]}&{vis.antenna_names[a2]}"
                        rowid = np.tile(vis.rowid[rows], stack)
                        for j, leaf_name in enumerate(leaf_names):
                            b = e * len(leaf_names) + j
                            k = int(rk[b])
                            leaf = Path(zarr_output_path) / "MAIN" / f"{outcolumn}" / name / leaf_name
                            write_svd_to_zarr((Uh[b, :, :k], Sh[b, :k], Vh[b, :k, :]), leaf, compressor, level, rowid)
                done += len(batch)
        return done

    ngpus = max(1, min(int(ngpus or 1), torch.cuda.device_count(), max(1, len(baselines))))
    if ngpus == 1:
        return run(get_engine(), baselines, data_dev)
    results, errors = [0] * ngpus, []

    def worker(g):
        try:
            off, cnt = shard_baselines(len(baselines), ngpus, g)
            eng = get_engine(g) if g == torch.cuda.current_device() else Engine(g)
            results[g] = run(eng, baselines[off:off + cnt], data_dev if (data_dev is not None and data_dev.device.index == g) else None)
        except BaseException as ex:          # re-raised in the caller's thread
            errors.append(ex)
    threads = [threading.Thread(target=worker, args=(g,)) for g in range(ngpus)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return sum(results)


def write_store_tables(zarr_path, vis, column="DATA", outcolumn="COMPRESSED_DATA"):
    """The MAIN / ANTENNA / POLARIZATION columns the decompressors need (the reference converts every MS table,
    compress_ms.py:138-194; the full MS -> zarr conversion is out of scope here). ROWID is a coordinate of MAIN, as in
    the dask-ms datasets the reference writes (decompress_ms.py:159 reads maintable.coords["ROWID"])."""
    import os

    from .zarr_leaf import write_group
    write_group(os.path.join(zarr_path, "MAIN"),
                {"ANTENNA1": (vis.antenna1, ("row",)), "ANTENNA2": (vis.antenna2, ("row",)), "ROWID": (vis.rowid, ("row",))},
                attrs={"visco_b200": {"column": column, "outcolumn": outcolumn}}, coordinates="ROWID")
    write_group(os.path.join(zarr_path, "ANTENNA"), {"NAME": (np.array(vis.antenna_names, dtype="U"), ("row",))})
    write_group(os.path.join(zarr_path, "POLARIZATION"),
                {"CORR_TYPE": (np.array([vis.corr_types], dtype=np.int32), ("row", "corr"))})


def write_store_flags(zarr_path, flags_packed, flags_row_packed):
    """Groups FLAGS and FLAGS_ROW holding np.packbits(FLAG) / np.packbits(FLAG_ROW) with a `row` coordinate (reference
    write_a_group_to_zarr, compress_ms.py:706-720; FLAGS_ROW is the reference's spelling, :482)."""
    import os

    from .zarr_leaf import write_group
    for group, p in (("FLAGS_ROW", flags_row_packed), ("FLAGS", flags_packed)):
        p = np.ascontiguousarray(p, dtype=np.uint8)
        write_group(os.path.join(zarr_path, group), {group: (p, ("row",)), "row": (np.arange(p.shape[0]), ("row",))})


def finalize_store(zarr_path, vis, chunk_size_row=10000, compressor="zstd", level=4, column="DATA"):
    """Root .zgroup + root .zmetadata with MAIN/ANTENNA/POLARIZATION/FLAGS/FLAGS_ROW and a chunk-less MAIN/DATA entry:
    the reference's decompressor takes DATA.shape / .dtype / .chunks and coords["ROWID"] from there
    (decompress_ms.py:151-161; the column's chunks are deleted after compression, compress_ms.py:934-939)."""
    from .zarr_leaf import consolidate_root, get_compressor, virtual_column_meta
    nrow, nchan, ncorr = vis.data.shape
    rows = int(min(max(1, int(chunk_size_row or nrow)), max(nrow, 1)))
    codec = get_compressor(compressor, level) if compressor else None
    virtual = virtual_column_meta("MAIN", "DATA", (nrow, nchan, ncorr), "<c8", (rows, nchan, ncorr), codec)
    if column and column != "DATA":
        virtual.update(virtual_column_meta("MAIN", column, (nrow, nchan, ncorr), "<c8", (rows, nchan, ncorr), codec))
    return consolidate_root(zarr_path, virtual)


def compress_full_ms(ms_path: str, zarr_path: str, consolidated: bool = True, chunk_size_row: int = 10000,
                     overwrite: bool = True, compressor: str = "zstd", level: int = 4, nworkers: int = 4,
                     nthreads: int = 2, memory_limit: str = "4GB", direct_to_workers: bool = True,
                     correlation: str = "XX,YY", correlation_optimized: bool = False, fieldid: int = 0, ddid: int = 0,
                     scan: int = 1, column: str = "DATA", outcolumn: str = "COMPRESSED_DATA", batch_size: int = 20,
                     dashboard_addr: str = None, host_addr: str = None, use_model_data: bool = False,
                     model_data: str = None, flag_estimate: bool = False, decorrelation: float = None,
                     compressionrank: int = None, flagvalue: int = None, antennas: list = None, ngpus: int = 1):
    """Same keyword arguments as the reference's compress_full_ms (compress_ms.py:782-811). The dask-cluster arguments
    (nworkers, nthreads, memory_limit, direct_to_workers, dashboard_addr, host_addr, chunk_size_row) are accepted and
    ignored: the batches run on the GPU(s) of this process - `ngpus` (additive keyword, default 1) spreads the baselines
    over that many GPUs of the box, one host thread and handle each (see compress_visdata). Flags are bit-packed into the FLAGS / FLAGS_ROW groups and
    flagged visibilities are replaced by model data or by a constant on the device (compress_ms.py:478-483, 530-562);
    the scipy-griddata estimator (flag_estimate, compress_ms.py:197-292) is out of scope and raises."""
    import os
    import shutil

    from .msdata import VisData
    from .zarr_leaf import write_group
    if not os.path.exists(ms_path):
        raise ValueError(f"MS path {ms_path} does not exist.")                     # reference :876-877
    if flag_estimate and not use_model_data:
        raise NotImplementedError("flag_estimate (scipy griddata interpolation) is out of scope of this build")
    if flagvalue and not use_model_data and isinstance(flagvalue, str):
        try:                                                                       # reference :549-558
            flagvalue = complex(flagvalue.replace(" ", ""))
        except ValueError:
            try:
                flagvalue = float(flagvalue)
            except ValueError:
                raise ValueError(f"Invalid flagvalue '{flagvalue}'. Use a float or complex format like '1+1j'.")
    if compressor is not None:
        from .zarr_leaf import get_compressor
        get_compressor(compressor, level)                                          # ValueError for unknown names (:51)
    vis = VisData.load(ms_path, column=column, scan=scan, fieldid=fieldid, ddid=ddid)
    if overwrite and os.path.exists(zarr_path):
        shutil.rmtree(zarr_path)                                                   # reference :95-96
    os.makedirs(zarr_path, exist_ok=True)
    # the few MAIN / ANTENNA / POLARIZATION columns the decompressor needs (the full MS -> zarr conversion is out of scope)
    nrow, nchan, ncorr = vis.data.shape
    write_store_tables(zarr_path, vis, column=column, outcolumn=outcolumn)
    import torch
    eng = get_engine()
    dev = f"cuda:{eng.device}"
    data_dev = torch.from_numpy(vis.data).to(dev)
    # flags: bit-packed groups FLAGS / FLAGS_ROW (reference :478-483; FLAGS_ROW is the reference's spelling)
    flag = vis.flag if vis.flag is not None else np.zeros(vis.data.shape, bool)
    flag_row = vis.flag_row if vis.flag_row is not None else np.zeros(nrow, bool)
    flag_dev = torch.from_numpy(flag).to(dev)
    write_store_flags(zarr_path, eng.packbits(flag_dev).cpu().numpy(), eng.packbits(torch.from_numpy(flag_row).to(dev)).cpu().numpy())
    # root consolidated metadata: what the reference's decompressor opens (decompress_ms.py:151-161, 240-246)
    finalize_store(zarr_path, vis, chunk_size_row=chunk_size_row, compressor=compressor, level=level, column=column)
    # WEIGHT_SPECTRUM: first correlation plane, rank 1, at the leaf <store>/WEIGHT_SPECTRUM (reference :486-503)
    if vis.weight_spectrum is not None:
        from .zarr_leaf import write_svd_to_zarr
        write_svd_to_zarr(weight_spectrum_rank1(vis.weight_spectrum), os.path.join(zarr_path, "WEIGHT_SPECTRUM"),
                          compressor, level, vis.rowid)
    # flagged-value replacement before the SVD (reference :530-566)
    if use_model_data:
        mod = vis.model_data
        if model_data is not None and model_data != "MODEL_DATA":
            mod = VisData.load(ms_path, column=model_data).data if str(ms_path).endswith(".npz") else None
        if mod is None:
            raise ValueError("use_model_data needs a MODEL_DATA column in the data set")
        eng.flag_replace(data_dev, flag_dev, model=torch.from_numpy(np.ascontiguousarray(mod, np.complex64)).to(dev))
    elif flagvalue:
        LOG.warning(f"Replacing flagged data with {flagvalue} - may amplify noise")      # reference :560
        eng.flag_replace(data_dev, flag_dev, value=complex(flagvalue))
    else:
        LOG.warning("No flag replacement specified - flagged data will not be replaced")  # reference :566
    return compress_visdata(vis, zarr_path, data_dev=data_dev, correlation=correlation, correlation_optimized=correlation_optimized,
                            decorrelation=decorrelation, compressionrank=compressionrank, outcolumn=outcolumn,
                            compressor=compressor, level=level, batch_size=batch_size, antennas=antennas, ngpus=ngpus)
