"""CPU: the on-disk contract of the hot path — zarr-v2 leaf stores U(time, mode), S(mode), WT(mode, channel)
(reference compress_ms.py:723-763 writer, decompress_ms.py:188-194 reader)."""
import gzip
import json
import os

import numpy as np
import pytest

from visco_b200 import zarr_leaf as zl


def _factors(m=36, k=3, n=16, seed=0):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))).astype(np.complex64)
    S = np.sort(rng.random(k).astype(np.float32))[::-1].copy()
    V = (rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))).astype(np.complex64)
    return U, S, V


@pytest.mark.parametrize("comp", ["zstd", "gzip", None])
def test_leaf_round_trip_and_layout(tmp_path, comp):
    U, S, V = _factors()
    rowid = np.arange(36) * 21 + 5
    leaf = tmp_path / "MAIN" / "COMPRESSED_DATA" / "ANT-0&ANT-1" / "XX"
    zl.write_svd_to_zarr((U, S, V), leaf, comp, 3, rowid)
    # layout the reference's reader (xr.open_zarr on the LEAF) needs
    assert json.loads((leaf / ".zgroup").read_text()) == {"zarr_format": 2}
    cons = json.loads((leaf / ".zmetadata").read_text())
    assert cons["zarr_consolidated_format"] == 1
    for name, dims, dt in (("U", ["time", "mode"], "<c8"), ("S", ["mode"], "<f4"), ("WT", ["mode", "channel"], "<c8"),
                           ("time", ["time"], "<i8"), ("mode", ["mode"], "<i8"), ("channel", ["channel"], "<i8")):
        za = json.loads((leaf / name / ".zarray").read_text())
        assert za["zarr_format"] == 2 and za["order"] == "C" and za["dtype"] == dt and za["filters"] is None
        assert json.loads((leaf / name / ".zattrs").read_text())["_ARRAY_DIMENSIONS"] == dims
        assert cons["metadata"][f"{name}/.zarray"] == za
        want = ({"id": comp, "level": 3} if comp else None) if name in ("U", "S", "WT") else None
        assert za["compressor"] == want
    assert json.loads((leaf / "U" / ".zarray").read_text())["shape"] == [36, 3]
    u, s, v, r = zl.read_svd_from_zarr(leaf)
    np.testing.assert_array_equal(u, U)
    np.testing.assert_array_equal(s, S)
    np.testing.assert_array_equal(v, V)
    np.testing.assert_array_equal(r, rowid)


def test_reader_accepts_a_foreign_chunking(tmp_path):
    """A leaf as another zarr-v2 writer (xarray + zarr 2.18) may lay it out: several chunks per array, nested chunk
    keys, gzip codec, partial edge chunks, a missing chunk (fill value), S stored as (k, 1)."""
    U, S, V = _factors(m=50, k=5, n=23, seed=3)
    leaf = tmp_path / "leaf"

    def put(name, a, chunks, dims, sep=".", codec={"id": "gzip", "level": 1}, skip=()):
        d = leaf / name
        d.mkdir(parents=True)
        meta = {"chunks": list(chunks), "compressor": codec, "dtype": a.dtype.str, "fill_value": None, "filters": None,
                "order": "C", "shape": list(a.shape), "zarr_format": 2}
        if sep == "/":
            meta["dimension_separator"] = "/"
        (d / ".zarray").write_text(json.dumps(meta))
        (d / ".zattrs").write_text(json.dumps({"_ARRAY_DIMENSIONS": dims}))
        grid = [range((s + c - 1) // c) for s, c in zip(a.shape, chunks)]
        import itertools
        for idx in itertools.product(*grid):
            if idx in skip:
                continue
            block = np.zeros(chunks, a.dtype)
            sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, a.shape))
            block[tuple(slice(0, x.stop - x.start) for x in sl)] = a[sl]
            p = d.joinpath(*[str(i) for i in idx]) if sep == "/" else d / ".".join(str(i) for i in idx)
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_bytes(gzip.compress(block.tobytes()) if codec else block.tobytes())

    leaf.mkdir()
    (leaf / ".zgroup").write_text('{"zarr_format": 2}')
    put("U", U, (16, 2), ["time", "mode"])
    put("S", S.reshape(-1, 1), (5, 1), ["mode", "one"], codec=None)
    put("WT", V, (2, 10), ["mode", "channel"], sep="/")
    u, s, v, r = zl.read_svd_from_zarr(leaf)
    np.testing.assert_array_equal(u, U)
    np.testing.assert_array_equal(s, S)
    np.testing.assert_array_equal(v, V)
    np.testing.assert_array_equal(r, np.arange(50))          # no time coordinate -> row numbers


def test_vlen_utf8_string_column(tmp_path):
    """ANTENNA/NAME as xarray writes string columns: object dtype + vlen-utf8 filter."""
    names = ["ANT-0", "m000", "ska-ü"]
    d = tmp_path / "NAME"
    d.mkdir()
    (d / ".zarray").write_text(json.dumps({"chunks": [2], "compressor": None, "dtype": "|O", "fill_value": None,
                                           "filters": [{"id": "vlen-utf8"}], "order": "C", "shape": [3], "zarr_format": 2}))
    for i, part in enumerate((names[:2], names[2:])):
        buf = np.array([len(part)], "<u4").tobytes()
        for s in part:
            b = s.encode()
            buf += np.array([len(b)], "<u4").tobytes() + b
        (d / str(i)).write_bytes(buf)
    assert list(zl.read_array(d)) == names


def test_codec_contract():
    assert zl.get_compressor(None, 3) is None
    assert zl.get_compressor("ZSTD", 4) == {"id": "zstd", "level": 4}
    assert zl.get_compressor("gzip", 2)["id"] == "gzip"
    with pytest.raises(ValueError):
        zl.get_compressor("lzma", 1)                          # reference compress_ms.py:51
    payload = os.urandom(1000) + bytes(5000)
    for codec in ({"id": "zstd", "level": 3}, {"id": "gzip", "level": 1}, {"id": "zlib", "level": 1}, None):
        assert zl._decode(zl._encode(payload, codec), codec, len(payload)) == payload


def test_writer_rejects_inconsistent_factors(tmp_path):
    U, S, V = _factors()
    with pytest.raises(ValueError):
        zl.write_svd_to_zarr((U, S[:2], V), tmp_path / "x", None, 1, np.arange(36))
    with pytest.raises(ValueError):
        zl.write_svd_to_zarr((U, S, V), tmp_path / "y", None, 1, np.arange(35))
    with pytest.raises(FileNotFoundError):
        zl.read_svd_from_zarr(tmp_path / "nothing")


def test_real_factors_are_stored_as_float32_and_bundle_keeps_weights(tmp_path):
    """The rank-1 WEIGHT_SPECTRUM leaf of the reference holds real float32 factors (compress_ms.py:493-498)."""
    import json

    from visco_b200.msdata import VisData
    from visco_b200.zarr_leaf import read_svd_from_zarr, write_svd_to_zarr
    rng = np.random.default_rng(0)
    u = rng.random((12, 1)).astype(np.float32)
    s = np.array([3.0], np.float32)
    v = rng.random((1, 5)).astype(np.float32)
    write_svd_to_zarr((u, s, v), tmp_path / "WEIGHT_SPECTRUM", "zstd", 3, np.arange(12))
    assert json.load(open(tmp_path / "WEIGHT_SPECTRUM" / "U" / ".zarray"))["dtype"] == "<f4"
    assert json.load(open(tmp_path / "WEIGHT_SPECTRUM" / "WT" / ".zarray"))["dtype"] == "<f4"
    U, S, WT, rowid = read_svd_from_zarr(tmp_path / "WEIGHT_SPECTRUM")
    np.testing.assert_array_equal(U.real, u)
    np.testing.assert_array_equal(WT.real, v)
    vis = VisData(data=np.zeros((6, 3, 4), np.complex64), antenna1=[0] * 6, antenna2=[1] * 6, antenna_names=["a", "b"],
                  weight_spectrum=rng.random((6, 3, 4)))
    vis.save(str(tmp_path / "b.npz"))
    back = VisData.load(str(tmp_path / "b.npz"))
    assert back.weight_spectrum.dtype == np.float32 and back.weight_spectrum.shape == (6, 3, 4)
    with pytest.raises(ValueError):
        VisData(data=np.zeros((6, 3, 4), np.complex64), antenna1=[0] * 6, antenna2=[1] * 6, antenna_names=["a", "b"],
                weight_spectrum=np.zeros((5, 3, 4)))
