"""CPU: the store-level on-disk contract — a store written by visco_b200 must be readable by the REFERENCE's decompressor.

xarray / zarr / dask-ms are not installed in this image (say so: the reference's decompressor itself cannot be run here
or on the GPU box), so this test walks exactly the accesses reference visco/decompress_ms.py makes, through the same
resolution rules zarr 2.x applies, against the zarr-v2 specification text:

  * ``xr.open_zarr(zarr_path, group=G, consolidated=True)`` (decompress_ms.py:151-152, 240, 245) -> zarr.open_consolidated
    reads ONLY ``<zarr_path>/.zmetadata`` (``{"zarr_consolidated_format": 1, "metadata": {key: json}}``); group G exists
    iff ``G/.zgroup`` is a key; its arrays are the keys ``G/<name>/.zarray``; xarray needs ``_ARRAY_DIMENSIONS`` in each
    ``G/<name>/.zattrs`` and promotes to coordinates every name listed in a ``coordinates`` attribute (CF) plus arrays
    named like a dimension. Chunks are then fetched from ``<zarr_path>/G/<name>/<i.j.k>``.
  * ``maintable.DATA.shape / .dtype / .chunks`` (decompress_ms.py:157-161) come from metadata alone: the column's chunks
    were deleted after compression (compress_ms.py:934-939).
  * ``xr.open_zarr(leaf)`` (decompress_ms.py:188-194) and ``xr.open_zarr(WEIGHT_SPECTRUM leaf, consolidated=True)``
    (:251) resolve against the ``.zmetadata`` at the root OF THE LEAF.
"""
import json
import os

import numpy as np
import pytest

from oracle import visco_oracle as vo
from visco_b200 import zarr_leaf as zl
from visco_b200.compress_ms import finalize_store, write_store_flags, write_store_tables
from visco_b200.msdata import VisData

ZARRAY_KEYS = {"zarr_format", "shape", "chunks", "dtype", "compressor", "fill_value", "order", "filters"}  # zarr v2 spec


class ConsolidatedGroup:
    """What xr.open_zarr(store, group=G, consolidated=True) can see: the root .zmetadata and the chunk files."""

    def __init__(self, store, group=None):
        self.store, self.group = store, group
        doc = json.load(open(os.path.join(store, ".zmetadata")))           # the ONLY metadata file consulted
        assert doc["zarr_consolidated_format"] == 1
        self.md = doc["metadata"]
        self.prefix = f"{group}/" if group else ""
        assert self.md[f"{self.prefix}.zgroup"] == {"zarr_format": 2}, f"group {group} missing from consolidated metadata"
        self.names = sorted({k[len(self.prefix):].split("/")[0] for k in self.md
                             if k.startswith(self.prefix) and k.endswith("/.zarray") and k[len(self.prefix):].count("/") == 1})
        self.coord_names = set()
        for n in self.names:
            za, at = self.zarray(n), self.zattrs(n)
            assert ZARRAY_KEYS <= set(za) and za["zarr_format"] == 2 and za["order"] in ("C", "F")
            assert len(za["shape"]) == len(za["chunks"]) == len(at["_ARRAY_DIMENSIONS"])   # KeyError = xarray cannot open it
            np.dtype(za["dtype"])
            self.coord_names |= set(str(at.get("coordinates", "")).split())
            if at["_ARRAY_DIMENSIONS"] == [n]:
                self.coord_names.add(n)

    def zarray(self, n):
        return self.md[f"{self.prefix}{n}/.zarray"]

    def zattrs(self, n):
        return self.md[f"{self.prefix}{n}/.zattrs"]

    def values(self, n):
        path = os.path.join(self.store, self.group or "", n)
        # the per-array files must agree with the consolidated copy (a non-consolidated open reads these)
        if os.path.exists(os.path.join(path, ".zarray")):
            assert json.load(open(os.path.join(path, ".zarray"))) == self.zarray(n)
        return zl.read_array(path)

    def coord(self, n):
        assert n in self.coord_names, f"{n} is not a coordinate of group {self.group}"
        return self.values(n)


def _toy_vis(nant=5, ntime=12, nchan=8, ncorr=4, seed=0):
    rng = np.random.default_rng(seed)
    a1, a2 = np.array([(i, j) for i in range(nant) for j in range(i, nant)]).T       # autocorrelations included
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    nrow = ant1.size
    data = (rng.standard_normal((nrow, nchan, ncorr)) + 1j * rng.standard_normal((nrow, nchan, ncorr))).astype(np.complex64)
    flag = rng.random((nrow, nchan, ncorr)) < 0.1
    return VisData(data=data, antenna1=ant1, antenna2=ant2, antenna_names=[f"ANT-{i}" for i in range(nant)],
                   rowid=np.arange(nrow) + 100, flag=flag, flag_row=flag.all(axis=(1, 2)))


def _write_store(tmp_path, vis, chunk_size_row=50):
    store = str(tmp_path / "out.zarr")
    os.makedirs(store)
    write_store_tables(store, vis)
    write_store_flags(store, np.packbits(vis.flag, axis=None), np.packbits(vis.flag_row, axis=None))
    leaves = {}
    for (a1, a2) in vis.baselines()[:3]:
        rows = vis.baseline_rows(a1, a2)
        name = f"{vis.antenna_names[a1]}&{vis.antenna_names[a2]}"
        for corr, ci in (("XX", 0), ("YY", 3)):
            u, s, vt = vo.ref_apply_svd(vis.data[rows][:, :, ci], compressionrank=3)
            leaf = os.path.join(store, "MAIN", "COMPRESSED_DATA", name, corr)
            zl.write_svd_to_zarr((u, s, vt), leaf, "zstd", 4, vis.rowid[rows])
            leaves[(name, corr)] = (u, s, vt, rows)
    w = np.abs(vis.data[:, :, 0]).astype(np.float32)
    uw, sw, vw = vo.ref_apply_svd(w, compressionrank=1)
    zl.write_svd_to_zarr((uw.real, sw, vw.real), os.path.join(store, "WEIGHT_SPECTRUM"), "zstd", 4, vis.rowid)
    finalize_store(store, vis, chunk_size_row=chunk_size_row, compressor="zstd", level=4)
    return store, leaves


def test_reference_decompressor_accesses_resolve_through_root_consolidated_metadata(tmp_path):
    vis = _toy_vis()
    store, leaves = _write_store(tmp_path, vis)
    nrow, nchan, ncorr = vis.data.shape
    # decompress_ms.py:151-161
    maintable = ConsolidatedGroup(store, "MAIN")
    antennas = ConsolidatedGroup(store, "ANTENNA")
    antnames = antennas.values("NAME")
    assert [str(x) for x in antnames] == vis.antenna_names
    np.testing.assert_array_equal(maintable.values("ANTENNA1"), vis.antenna1)
    np.testing.assert_array_equal(maintable.values("ANTENNA2"), vis.antenna2)
    za = maintable.zarray("DATA")
    assert za["shape"] == [nrow, nchan, ncorr] and np.dtype(za["dtype"]) == np.complex64          # .shape, .dtype
    assert za["chunks"] == [50, nchan, ncorr] and za["compressor"] == {"id": "zstd", "level": 4}    # .chunks
    assert maintable.zattrs("DATA")["_ARRAY_DIMENSIONS"] == ["row", "chan", "corr"]
    assert not os.path.exists(os.path.join(store, "MAIN", "DATA"))      # metadata only, like the reference after :936
    np.testing.assert_array_equal(maintable.coord("ROWID"), vis.rowid)  # maintable.coords["ROWID"]
    assert "ROWID" in maintable.zattrs("ANTENNA1")["coordinates"].split()
    # decompress_ms.py:163-199: leaf discovery by directory names, antenna names -> indices, leaf arrays
    base = os.path.join(store, "MAIN", "COMPRESSED_DATA")
    baselines = zl.list_subtables(base)
    assert set(baselines) == {k[0] for k in leaves}
    out = np.zeros((nrow, nchan, ncorr), np.complex64)
    corr_indices = {"XX": 0, "XY": 1, "YX": 2, "YY": -1}
    for baseline in baselines:
        a1n, a2n = baseline.split("&")
        i1, i2 = np.where(antnames == a1n)[0][0], np.where(antnames == a2n)[0][0]
        row_indices = np.where((maintable.values("ANTENNA1") == i1) & (maintable.values("ANTENNA2") == i2))[0]
        for corr_name in zl.list_subtables(os.path.join(base, baseline)):
            leaf = ConsolidatedGroup(os.path.join(base, baseline, corr_name))      # xr.open_zarr(leaf): leaf-rooted metadata
            U, S, Vt = leaf.values("U"), leaf.values("S"), leaf.values("WT")
            assert leaf.zattrs("U")["_ARRAY_DIMENSIONS"] == ["time", "mode"]
            assert {"time", "mode", "channel"} <= leaf.coord_names
            np.testing.assert_array_equal(leaf.coord("time"), vis.rowid[row_indices])
            out[row_indices, :, corr_indices[corr_name]] = vo.ref_reconstruct_vis(U, S, Vt)   # reference :216-232
            u, s, vt, rows = leaves[(baseline, corr_name)]
            np.testing.assert_array_equal(rows, row_indices)
            np.testing.assert_array_equal(U, u)
    # decompress_ms.py:240-246
    flags_ds = ConsolidatedGroup(store, "FLAGS")
    flags = np.unpackbits(flags_ds.values("FLAGS"), count=nrow * nchan * ncorr).reshape(nrow, nchan, ncorr)
    np.testing.assert_array_equal(flags.astype(bool), vis.flag)
    assert "row" in flags_ds.coord_names
    flag_row_ds = ConsolidatedGroup(store, "FLAGS_ROW")
    np.testing.assert_array_equal(np.unpackbits(flag_row_ds.values("FLAGS_ROW"), count=nrow).astype(bool), vis.flag_row)
    # decompress_ms.py:248-254
    assert "WEIGHT_SPECTRUM" in zl.list_subtables(store)
    weights = ConsolidatedGroup(os.path.join(store, "WEIGHT_SPECTRUM"))
    wrec = np.dot(weights.values("U"), np.diag(weights.values("S")))
    assert wrec.shape == (nrow, 1)
    # POLARIZATION is what the compressor itself re-opens (compress_ms.py:453)
    np.testing.assert_array_equal(ConsolidatedGroup(store, "POLARIZATION").values("CORR_TYPE"), [[9, 10, 11, 12]])


def test_root_metadata_follows_the_zarr_v2_consolidated_format(tmp_path):
    vis = _toy_vis(nant=3, ntime=4)
    store, _ = _write_store(tmp_path, vis, chunk_size_row=10 ** 6)
    assert json.load(open(os.path.join(store, ".zgroup"))) == {"zarr_format": 2}
    doc = json.load(open(os.path.join(store, ".zmetadata")))
    assert set(doc) == {"metadata", "zarr_consolidated_format"}
    md = doc["metadata"]
    assert md[".zgroup"] == {"zarr_format": 2}
    for g in ("MAIN", "ANTENNA", "POLARIZATION", "FLAGS", "FLAGS_ROW"):
        assert md[f"{g}/.zgroup"] == {"zarr_format": 2} and isinstance(md[f"{g}/.zattrs"], dict)
    # every key is <path>/.zgroup | .zattrs | .zarray, no chunk keys, no leaf trees
    assert all(k.rsplit("/", 1)[-1] in (".zgroup", ".zattrs", ".zarray") for k in md)
    assert not any(k.startswith("MAIN/COMPRESSED_DATA") for k in md)
    assert md["MAIN/DATA/.zarray"]["chunks"][0] == vis.data.shape[0]       # row chunk never exceeds the table
    assert md["MAIN/DATA/.zarray"]["fill_value"] is None


def test_our_reader_uses_the_same_contract(tmp_path):
    """visco_b200.decompress_ms._store_index takes the column shape from the root consolidated metadata as well (no
    private attribute), so stores from either implementation are read the same way."""
    from visco_b200.decompress_ms import _store_index, leaf_planes
    vis = _toy_vis(nant=3, ntime=4)
    store, _ = _write_store(tmp_path, vis)
    ant1, ant2, rowid, names, shape = _store_index(store)
    assert list(shape) == list(vis.data.shape) and names == vis.antenna_names
    np.testing.assert_array_equal(rowid, vis.rowid)
    assert "data_shape" not in json.load(open(os.path.join(store, "MAIN", ".zattrs"))).get("visco_b200", {})
    assert leaf_planes(store, 4)["YY"] == (3,) and leaf_planes(store, 4)["offdiagonals"] == (1, 2)
