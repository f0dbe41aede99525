"""GPU: the reference's own smoke configuration (tests/compression_tests.py:10-32 — XX,YY, decorrelation 0.90,
batch_size 10, zstd level 3) through the drop-in drivers on the sample-MS bundle, checked against the oracle; plus
cross-reading of leaf stores in both directions."""
import os

import numpy as np
import pytest

from oracle import visco_oracle as vo
from visco_b200.msdata import VisData

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUNDLE = os.path.join(ROOT, "tests", "golden", "sample_ms_kat7.npz")
KW = dict(consolidated=True, chunk_size_row=5000, overwrite=True, compressor="zstd", level=3, nworkers=1, nthreads=1,
          memory_limit="2GB", direct_to_workers=False, fieldid=0, ddid=0, scan=1, column="DATA",
          outcolumn="COMPRESSED_DATA", batch_size=10, dashboard_addr=None, host_addr=None)


def _oracle_decompressed(vis, correlation, optimized, **rank):
    out = np.zeros_like(vis.data)
    corr_idx = {"XX": 0, "XY": 1, "YX": 2, "YY": 3}
    for a1, a2 in vis.baselines():
        rows = vis.baseline_rows(a1, a2)
        blk = vis.data[rows]
        if optimized:
            for pair in (("XX", "YY"), ("XY", "YX")):
                m = np.vstack([blk[:, :, corr_idx[pair[0]]], blk[:, :, corr_idx[pair[1]]]])
                rec, _, _ = vo.roundtrip(m, **rank)
                parts = vo.ref_unstack_vis(rec, len(rows))
                out[rows, :, corr_idx[pair[0]]] = parts[0]
                out[rows, :, corr_idx[pair[1]]] = parts[1]
        else:
            for c in correlation.split(","):
                rec, _, _ = vo.roundtrip(blk[:, :, corr_idx[c]], **rank)
                out[rows, :, corr_idx[c]] = rec
    return out


@pytest.mark.parametrize("correlation,optimized,rank", [
    ("XX,YY", False, dict(decorrelation=0.90)),               # the reference's own test configuration
    ("XX,XY,YX,YY", False, dict(compressionrank=2)),
    ("XX,XY,YX,YY", True, dict(compressionrank=3)),           # --correlation-optimized: stacked leaves
])
def test_compress_decompress_sample_ms(tmp_path, correlation, optimized, rank):
    from visco_b200.compress_ms import compress_full_ms
    from visco_b200.decompress_ms import open_dataset, write_datasets_to_ms
    from visco_b200.zarr_leaf import list_subtables, read_svd_from_zarr
    vis = VisData.load(BUNDLE)
    z = str(tmp_path / "store.zarr")
    n = compress_full_ms(ms_path=BUNDLE, zarr_path=z, correlation=correlation, correlation_optimized=optimized, **KW, **rank)
    assert n == 6
    base = os.path.join(z, "MAIN", "COMPRESSED_DATA")
    assert list_subtables(base) == ["ANT-0&ANT-1", "ANT-0&ANT-2", "ANT-0&ANT-3", "ANT-1&ANT-2", "ANT-1&ANT-3", "ANT-2&ANT-3"]
    leaves = list_subtables(os.path.join(base, "ANT-0&ANT-1"))
    assert leaves == (["diagonals", "offdiagonals"] if optimized else sorted(correlation.split(",")))
    U, S, WT, rowid = read_svd_from_zarr(os.path.join(base, "ANT-1&ANT-3", leaves[0]))
    assert U.shape[0] == (720 if optimized else 360) and WT.shape[1] == 16 and U.shape[1] == len(S) == WT.shape[0]
    np.testing.assert_array_equal(rowid[:360], vis.rowid[vis.baseline_rows(1, 3)])
    # decompress through the driver and through the bundle writer
    out = open_dataset(z, "COMPRESSED_DATA", 7)
    ms_out = write_datasets_to_ms(z, str(tmp_path / "decompressed.npz"), "COMPRESSED_DATA", 50)
    np.testing.assert_array_equal(VisData.load(ms_out).data, out.data)
    want = _oracle_decompressed(vis, correlation, optimized, **rank)
    scale = np.abs(vis.data).max()
    assert np.abs(out.data - want).max() <= 5e-5 * scale, np.abs(out.data - want).max() / scale
    if not optimized and correlation == "XX,YY":
        assert not out.data[:, :, 1].any() and not out.data[:, :, 2].any()      # uncompressed correlations stay zero


def test_cross_reading_both_directions(tmp_path):
    """Leaves written from the REFERENCE's factors (oracle = its restatement) are read and reconstructed by us; leaves
    we write are reconstructed by the reference's arithmetic."""
    from visco_b200.compress_ms import apply_svd
    from visco_b200.decompress_ms import reconstruct_vis
    from visco_b200.zarr_leaf import read_svd_from_zarr, write_svd_to_zarr
    vis = VisData.load(BUNDLE)
    a = vis.data[vis.baseline_rows(0, 2)][:, :, 0]
    u, s, vt = vo.ref_apply_svd(a, compressionrank=4)
    write_svd_to_zarr((u, s, vt), tmp_path / "ref_leaf", "zstd", 4, vis.rowid[vis.baseline_rows(0, 2)])
    U, S, WT, _ = read_svd_from_zarr(tmp_path / "ref_leaf")
    ours = reconstruct_vis(U, S, WT)
    np.testing.assert_allclose(ours, vo.ref_reconstruct_vis(u, s, vt), atol=3e-5 * np.abs(a).max())
    write_svd_to_zarr(apply_svd(a, compressionrank=4), tmp_path / "our_leaf", "gzip", 2, vis.rowid[vis.baseline_rows(0, 2)])
    U, S, WT, _ = read_svd_from_zarr(tmp_path / "our_leaf")
    theirs = vo.ref_reconstruct_vis(U, S, WT)
    e_ref = np.linalg.norm(a - vo.ref_reconstruct_vis(u, s, vt))
    assert abs(np.linalg.norm(a - theirs) - e_ref) <= 1e-5 * e_ref + 5e-6 * np.linalg.norm(a)


def test_gather_scatter_kernels_match_numpy_indexing():
    """vk_gather_baselines / vk_scatter_baselines against the reference's slicing (compress_ms.py:591-664) and scatter
    (decompress_ms.py:216-232), bit-exact (pure data movement), with --correlation-optimized stacking and padding."""
    import torch
    from visco_b200.engine import get_engine
    eng = get_engine(0)
    vis = VisData.load(BUNDLE)
    data = torch.from_numpy(vis.data).to("cuda:0")
    bls = vis.baselines()
    rows = np.stack([vis.baseline_rows(a, b) for a, b in bls]).astype(np.int32)       # [6, 360]
    rows[2, 300:] = -1                                                                  # ragged entry: padded tail
    row_idx = torch.from_numpy(rows).to("cuda:0")
    # plain: XX, YY of every baseline
    sel = torch.tensor([[0, 3]] * 6, dtype=torch.int32, device="cuda:0")
    cube = eng.gather_baselines(data, row_idx, sel, 1).cpu().numpy()
    for e in range(6):
        valid = rows[e] >= 0
        np.testing.assert_array_equal(cube[2 * e][valid], vis.data[rows[e][valid]][:, :, 0])
        np.testing.assert_array_equal(cube[2 * e + 1][valid], vis.data[rows[e][valid]][:, :, 3])
        assert not cube[2 * e][~valid].any()
    # stacked: diagonals (XX over YY) and offdiagonals (XY over YX)
    sel4 = torch.tensor([[0, 3, 1, 2]] * 6, dtype=torch.int32, device="cuda:0")
    cube2 = eng.gather_baselines(data, row_idx, sel4, 2).cpu().numpy()
    assert cube2.shape == (12, 720, 16)
    np.testing.assert_array_equal(cube2[0], np.vstack([vis.data[rows[0]][:, :, 0], vis.data[rows[0]][:, :, 3]]))
    np.testing.assert_array_equal(cube2[3], np.vstack([vis.data[rows[1]][:, :, 1], vis.data[rows[1]][:, :, 2]]))
    # scatter is the inverse on the selected entries and leaves everything else untouched
    out = torch.full_like(data, 7.0)
    eng.scatter_baselines(torch.from_numpy(cube2).to("cuda:0"), out, row_idx, sel4, 2)
    out = out.cpu().numpy()
    touched = np.zeros(vis.data.shape[0], bool)
    for e in range(6):
        touched[rows[e][rows[e] >= 0]] = True
    np.testing.assert_array_equal(out[touched], vis.data[touched])
    assert np.all(out[~touched] == 7.0)


def test_flag_kernels_and_flag_replacement_end_to_end(tmp_path):
    """next-3: np.packbits / np.unpackbits on device are bit-exact; flagged visibilities are replaced (constant or
    model column) before the SVD and the flags survive the round trip (reference compress_ms.py:478-483, 530-562;
    decompress_ms.py:240-246)."""
    import torch
    from visco_b200.compress_ms import compress_full_ms
    from visco_b200.decompress_ms import open_dataset
    from visco_b200.engine import get_engine
    eng = get_engine(0)
    rng = np.random.default_rng(7)
    for n in (1, 7, 8, 9, 1000, 138240 + 3):
        f = rng.random(n) < 0.3
        packed = eng.packbits(torch.from_numpy(f).to("cuda:0")).cpu().numpy()
        np.testing.assert_array_equal(packed, np.packbits(f, axis=None))
        back = eng.unpackbits(torch.from_numpy(packed).to("cuda:0"), n).cpu().numpy()
        np.testing.assert_array_equal(back, np.unpackbits(packed, count=n))
    vis = VisData.load(BUNDLE)
    flag = rng.random(vis.data.shape) < 0.02
    flag_row = rng.random(vis.data.shape[0]) < 0.01
    model = (vis.data * 0.5).astype(np.complex64)
    dirty = vis.data.copy()
    dirty[flag] = 1e4                                            # RFI-like outliers under the flags
    bundle = str(tmp_path / "flagged.npz")
    VisData(data=dirty, antenna1=vis.antenna1, antenna2=vis.antenna2, antenna_names=vis.antenna_names, rowid=vis.rowid,
            flag=flag, flag_row=flag_row, model_data=model).save(bundle)
    for kw, clean in ((dict(flagvalue="1+1j"), np.where(flag, 1 + 1j, dirty)), (dict(use_model_data=True), np.where(flag, model, dirty))):
        z = str(tmp_path / "flagged.zarr")
        compress_full_ms(ms_path=bundle, zarr_path=z, correlation="XX,YY", correlation_optimized=False, compressionrank=16,
                         **KW, **kw)
        out = open_dataset(z, "COMPRESSED_DATA", 50)
        np.testing.assert_array_equal(out.flag, flag)
        np.testing.assert_array_equal(out.flag_row, flag_row)
        # full rank (16 channels): the decompressed column equals the flag-replaced input, not the dirty one
        for c in (0, 3):
            np.testing.assert_allclose(out.data[:, :, c], clean[:, :, c].astype(np.complex64), atol=2e-4 * np.abs(clean).max())


def test_weight_spectrum_rank1_and_the_reference_decompression_quirk(tmp_path):
    """WEIGHT_SPECTRUM[:, :, 0] is stored as its leading singular triplet (compress_ms.py:489-503) in
    float32, with dask's svd_flip sign; decompression returns U.diag(S) only, tiled over the correlations, for both
    WEIGHT_SPECTRUM and SIGMA_SPECTRUM (decompress_ms.py:248-270)."""
    import json

    from visco_b200.compress_ms import compress_full_ms, weight_spectrum_rank1
    from visco_b200.decompress_ms import open_dataset
    from visco_b200.zarr_leaf import read_svd_from_zarr
    vis = VisData.load(BUNDLE)
    nrow, nchan, ncorr = vis.data.shape
    rng = np.random.default_rng(7)
    ws = (1.0 + rng.random((nrow, 1, 1))) * (0.5 + rng.random((1, nchan, 1))) * (1 + 0.05 * rng.random((nrow, nchan, ncorr)))
    vis.weight_spectrum = ws.astype(np.float32)
    bundle = str(tmp_path / "with_weights.npz")
    vis.save(bundle)
    assert VisData.load(bundle).weight_spectrum.shape == (nrow, nchan, ncorr)
    store = str(tmp_path / "w.zarr")
    compress_full_ms(bundle, store, correlation="XX,YY", compressionrank=2, **KW)
    leaf = os.path.join(store, "WEIGHT_SPECTRUM")
    assert json.load(open(os.path.join(leaf, "U", ".zarray")))["dtype"] == "<f4"       # real factors stay real
    U, S, WT, rowid = read_svd_from_zarr(leaf)
    assert U.shape == (nrow, 1) and S.shape == (1,) and WT.shape == (1, nchan)
    np.testing.assert_array_equal(rowid, vis.rowid)
    u, s, vt = vo.ref_apply_svd(vis.weight_spectrum[:, :, 0], compressionrank=1)        # the reference, float32 LAPACK
    np.testing.assert_allclose(S, s, rtol=1e-4)
    np.testing.assert_allclose(U.real, u, atol=2e-5 * np.abs(u).max())                  # same sign convention
    np.testing.assert_allclose(WT.real, vt, atol=2e-5 * np.abs(vt).max())
    assert not U.imag.any() and not WT.imag.any()
    u1, s1, v1 = weight_spectrum_rank1(vis.weight_spectrum)
    assert u1.dtype == np.float32 and v1.dtype == np.float32 and v1.sum() >= 0
    out = open_dataset(store)
    ref = np.tile(np.expand_dims(np.dot(u, np.diag(s)), axis=-1), (1, 1, ncorr))        # decompress_ms.py:252-254
    assert out.weight_spectrum.shape == (nrow, 1, ncorr)
    np.testing.assert_allclose(out.weight_spectrum, ref, rtol=1e-4, atol=1e-6)
    assert out.sigma_spectrum is out.weight_spectrum or np.array_equal(out.sigma_spectrum, out.weight_spectrum)


# ---------------------------------------------------------------------------------------------------------------------
# The sample Measurement Set itself through the command line (reference tests/compression_tests.py:5-32,
# decompression_tests.py:10-12): every one of its 21 baselines, read by visco_b200.casatable (no python-casacore here).
# ---------------------------------------------------------------------------------------------------------------------
SAMPLE_MS = os.path.join(ROOT, "tests", "golden", "sim-visco-kat7-subset.ms")


def test_cli_compresses_the_sample_measurement_set(tmp_path):
    from click.testing import CliRunner
    from visco_b200.decompress_ms import open_dataset
    from visco_b200.parser_config import cli
    from visco_b200.zarr_leaf import list_subtables
    z = str(tmp_path / "sim-visco-kat7.zarr")
    r = CliRunner().invoke(cli, ["compressms", "-ms", SAMPLE_MS, "-zs", z, "-corr", "XX,YY", "-dec", "0.90", "-bs", "10",
                                 "-l", "3", "-nw", "1", "-nt", "1"], catch_exceptions=False)
    assert r.exit_code == 0, r.output
    base = os.path.join(z, "MAIN", "COMPRESSED_DATA")
    assert len(list_subtables(base)) == 21 and list_subtables(os.path.join(base, "ANT-2&ANT-5")) == ["XX", "YY"]
    assert os.path.exists(os.path.join(z, ".zmetadata"))
    vis = VisData.load(SAMPLE_MS)
    back = open_dataset(z)
    ref = _oracle_decompressed(vis, "XX,YY", False, decorrelation=0.90)
    for c in (0, 3):
        num = np.linalg.norm(back.data[:, :, c] - ref[:, :, c])
        assert num <= 2e-4 * np.linalg.norm(ref[:, :, c]), c       # same ranks, same projections as the reference's path
    assert not back.data[:, :, 1:3].any() and back.flag.shape == vis.data.shape and not back.flag.any()
    out = str(tmp_path / "decompressed.npz")
    r = CliRunner().invoke(cli, ["decompressms", "-zs", z, "-ms", out], catch_exceptions=False)
    assert r.exit_code == 0 and VisData.load(out).data.shape == vis.data.shape


def test_sharded_drivers_match_the_single_gpu_result(tmp_path):
    """--ngpus: baselines / reconstruction tasks split over the GPUs of the box, one host thread and handle each
    (the role of the reference's nworkers, compress_ms.py:571-697). Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from visco_b200.compress_ms import compress_full_ms
    from visco_b200.decompress_ms import open_dataset
    from visco_b200.zarr_leaf import list_subtables, read_svd_from_zarr
    z1, z2 = str(tmp_path / "one.zarr"), str(tmp_path / "two.zarr")
    kw = dict(KW, correlation="XX,YY", compressionrank=3)
    assert compress_full_ms(ms_path=SAMPLE_MS, zarr_path=z1, **kw) == 21
    assert compress_full_ms(ms_path=SAMPLE_MS, zarr_path=z2, ngpus=2, **kw) == 21
    b1, b2 = (os.path.join(z, "MAIN", "COMPRESSED_DATA") for z in (z1, z2))
    assert list_subtables(b1) == list_subtables(b2)
    for bl in list_subtables(b1):
        for c in ("XX", "YY"):
            f1, f2 = read_svd_from_zarr(os.path.join(b1, bl, c)), read_svd_from_zarr(os.path.join(b2, bl, c))
            np.testing.assert_array_equal(f1[1], f2[1])              # same kernels, same inputs: identical singular values
    d1, d2 = open_dataset(z1), open_dataset(z2, ngpus=2)
    np.testing.assert_array_equal(d1.data, d2.data)
