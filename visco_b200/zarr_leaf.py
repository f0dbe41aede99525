"""Zarr-v2 leaf stores for the compressed factors — the on-disk contract of the hot path.

The reference writes one zarr ``DirectoryStore`` PER (baseline, correlation) leaf with ``xarray.Dataset.to_zarr``
(reference visco/compress_ms.py:723-763): data variables ``U(time, mode)``, ``S(mode)``, ``WT(mode, channel)`` and
coordinates ``time = ROWID``, ``mode = arange(k)``, ``channel = arange(nchan)``, each with the selected numcodecs
compressor; and reads them back with ``xr.open_zarr(leaf)`` using only the ``U``, ``S`` and ``WT`` arrays
(reference visco/decompress_ms.py:188-194).

zarr / xarray / numcodecs are not installed here, so this module writes and parses the zarr-v2 layout directly:
``.zgroup``, ``.zattrs``, consolidated ``.zmetadata``, one directory per array with ``.zarray`` + ``.zattrs``
(``_ARRAY_DIMENSIONS`` is what xarray needs to rebuild the dataset) and C-order chunk files named ``i.j``.
Codecs: ``zstd`` through ctypes -> libzstd.so.1 (numcodecs id "zstd"), ``gzip`` through the stdlib (id "gzip"),
``None`` -> uncompressed. ``blosc`` needs a blosc library and raises if none is importable.
The writer emits one chunk per array (always valid); the reader accepts any regular chunk grid, "." or "/" chunk-key
separators, little- or big-endian dtypes, and missing chunks (fill_value), i.e. whatever xarray chose to write.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import gzip as _gzip
import itertools
import json
import os
import zlib
from pathlib import Path

import numpy as np

# ------------------------------------------------------------------------------------------------- codecs
_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        lib = ctypes.CDLL(name)
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _zstd = lib
    return _zstd


def get_compressor(name: str = None, level: int = None):
    """Same contract as the reference's get_compressor (compress_ms.py:33-51): returns a codec description (the JSON
    that goes into ``.zarray``) or None; unknown names raise ValueError."""
    if name is None:
        return None
    low = name.lower()
    if low == "zstd":
        return {"id": "zstd", "level": int(level if level is not None else 1)}
    if low == "gzip":
        return {"id": "gzip", "level": int(level if level is not None else 1)}
    if low == "blosc":
        return {"id": "blosc", "cname": "lz4", "clevel": int(level if level is not None else 5), "shuffle": 1, "blocksize": 0}
    raise ValueError(f"Unsupported compressor: {name}")


def _encode(buf: bytes, codec) -> bytes:
    if codec is None:
        return buf
    cid = codec["id"]
    if cid == "zstd":
        lib = _libzstd()
        cap = lib.ZSTD_compressBound(len(buf))
        out = ctypes.create_string_buffer(cap)
        n = lib.ZSTD_compress(out, cap, buf, len(buf), int(codec.get("level", 1)))
        if lib.ZSTD_isError(n):
            raise RuntimeError("ZSTD_compress failed")
        return out.raw[:n]
    if cid == "gzip":
        return _gzip.compress(buf, compresslevel=int(codec.get("level", 1)), mtime=0)
    if cid == "zlib":
        return zlib.compress(buf, int(codec.get("level", 1)))
    if cid == "blosc":
        try:
            import blosc  # noqa: F401
        except ImportError as e:
            raise RuntimeError("compressor 'blosc' needs the python-blosc package, which is not installed") from e
        import blosc
        return blosc.compress(buf, typesize=8, clevel=int(codec.get("clevel", 5)), cname=codec.get("cname", "lz4"),
                              shuffle=int(codec.get("shuffle", 1)))
    raise ValueError(f"unsupported codec {cid}")


def _decode(buf: bytes, codec, nbytes: int) -> bytes:
    if codec is None:
        return buf
    cid = codec["id"]
    if cid == "zstd":
        lib = _libzstd()
        size = lib.ZSTD_getFrameContentSize(buf, len(buf))
        cap = int(size) if size not in (2 ** 64 - 1, 2 ** 64 - 2) else nbytes
        out = ctypes.create_string_buffer(max(cap, 1))
        n = lib.ZSTD_decompress(out, cap, buf, len(buf))
        if lib.ZSTD_isError(n):
            raise RuntimeError("ZSTD_decompress failed (corrupt chunk?)")
        return out.raw[:n]
    if cid == "gzip":
        return _gzip.decompress(buf)
    if cid == "zlib":
        return zlib.decompress(buf)
    if cid == "blosc":
        try:
            import blosc
        except ImportError as e:
            raise RuntimeError("this store was written with blosc; python-blosc is not installed") from e
        return blosc.decompress(buf)
    raise ValueError(f"unsupported codec {cid}")


# ------------------------------------------------------------------------------------------------- arrays
def _fill_json(dtype: np.dtype):
    # xarray turns a zarr-v2 fill_value into `_FillValue` and masks equal data values when it decodes a store, so the
    # only safe choices are NaN for floats (what xarray itself writes; NaN never occurs in factors) and null elsewhere.
    if dtype.kind == "f":
        return "NaN"
    if dtype.kind == "U":
        return ""
    return None


def _zarray_meta(a: np.ndarray, codec) -> dict:
    return {"chunks": [int(s) for s in a.shape] if a.ndim else [], "compressor": codec, "dtype": a.dtype.str,
            "fill_value": _fill_json(a.dtype), "filters": None, "order": "C", "shape": [int(s) for s in a.shape],
            "zarr_format": 2}


def write_array(group_dir: Path, name: str, a: np.ndarray, dims, codec=None, extra_attrs: dict | None = None) -> dict:
    """Write one zarr-v2 array as a single chunk; returns {relative key: json} for the consolidated metadata."""
    a = np.ascontiguousarray(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    d = Path(group_dir) / name
    d.mkdir(parents=True, exist_ok=True)
    meta = _zarray_meta(a, codec)
    attrs = {"_ARRAY_DIMENSIONS": list(dims), **(extra_attrs or {})}
    (d / ".zarray").write_text(json.dumps(meta, indent=4, sort_keys=True))
    (d / ".zattrs").write_text(json.dumps(attrs, indent=4))
    if a.size:
        key = ".".join("0" for _ in a.shape) if a.ndim else "0"
        (d / key).write_bytes(_encode(a.tobytes(), codec))
    return {f"{name}/.zarray": meta, f"{name}/.zattrs": attrs}


def write_group(group_dir: Path, arrays: dict, attrs: dict | None = None, codec_for=None, coordinates: str | None = None):
    """arrays: name -> (ndarray, dims). Writes .zgroup/.zattrs/.zmetadata and every array. `coordinates` names a
    non-dimension coordinate array of the group (e.g. "ROWID"): it is recorded CF-style in the ``coordinates`` attribute
    of every other array that shares its dimension, which is how xarray marks it as a coordinate on read."""
    group_dir = Path(group_dir)
    group_dir.mkdir(parents=True, exist_ok=True)
    zgroup = {"zarr_format": 2}
    (group_dir / ".zgroup").write_text(json.dumps(zgroup, indent=4))
    (group_dir / ".zattrs").write_text(json.dumps(attrs or {}, indent=4))
    cons = {".zattrs": attrs or {}, ".zgroup": zgroup}
    cdims = set(arrays[coordinates][1]) if coordinates and coordinates in arrays else set()
    for name, (a, dims) in arrays.items():
        extra = {"coordinates": coordinates} if (cdims and name != coordinates and cdims <= set(dims)) else None
        cons.update(write_array(group_dir, name, a, dims, codec_for(name) if codec_for else None, extra))
    (group_dir / ".zmetadata").write_text(json.dumps({"metadata": cons, "zarr_consolidated_format": 1}, indent=4))


def consolidate_root(zarr_path, virtual: dict | None = None):
    """Root ``.zgroup`` + root ``.zmetadata`` of a compressed store: the consolidated view the reference's decompressor
    opens (``xr.open_zarr(zarr_path, group="MAIN" | "ANTENNA" | "FLAGS" | "FLAGS_ROW", consolidated=True)``,
    reference decompress_ms.py:151-152, 240, 245 - zarr resolves ``consolidated=True`` against the ``.zmetadata`` at the
    ROOT of the store and never looks at per-group files). Keys are ``<group>/.zgroup``, ``<group>/.zattrs``,
    ``<group>/<array>/.zarray``, ``<group>/<array>/.zattrs`` for every top-level group and the arrays directly inside
    it (the per-baseline leaf trees below MAIN/<outcolumn> are separate stores with their own ``.zmetadata``, as in the
    reference, and are not listed). ``virtual`` adds metadata-only entries: the reference deletes the chunks of the
    raw visibility column but still reads ``maintable.DATA.shape / .dtype / .chunks`` from this file
    (compress_ms.py:934-939 vs decompress_ms.py:157-161)."""
    root = Path(zarr_path)
    zgroup = {"zarr_format": 2}
    (root / ".zgroup").write_text(json.dumps(zgroup, indent=4))
    md = {".zgroup": zgroup}
    for g in sorted(p for p in root.iterdir() if p.is_dir() and (p / ".zgroup").exists()):
        md[f"{g.name}/.zgroup"] = json.loads((g / ".zgroup").read_text())
        md[f"{g.name}/.zattrs"] = json.loads((g / ".zattrs").read_text()) if (g / ".zattrs").exists() else {}
        for a in sorted(p for p in g.iterdir() if p.is_dir() and (p / ".zarray").exists()):
            md[f"{g.name}/{a.name}/.zarray"] = json.loads((a / ".zarray").read_text())
            md[f"{g.name}/{a.name}/.zattrs"] = json.loads((a / ".zattrs").read_text()) if (a / ".zattrs").exists() else {}
    md.update(virtual or {})
    (root / ".zmetadata").write_text(json.dumps({"metadata": md, "zarr_consolidated_format": 1}, indent=4))
    return md


def virtual_column_meta(group: str, name: str, shape, dtype: str, chunks, codec=None, dims=("row", "chan", "corr"),
                        coordinates: str = "ROWID") -> dict:
    """``.zarray`` / ``.zattrs`` entries of a column that has metadata but no chunks (see consolidate_root)."""
    dt = np.dtype(dtype)
    return {f"{group}/{name}/.zarray": {"chunks": [int(c) for c in chunks], "compressor": codec, "dtype": dt.str,
                                         "fill_value": _fill_json(dt), "filters": None, "order": "C",
                                         "shape": [int(x) for x in shape], "zarr_format": 2},
            f"{group}/{name}/.zattrs": {"_ARRAY_DIMENSIONS": list(dims), "coordinates": coordinates}}


def _read_vlen_utf8(array_dir: Path, meta: dict) -> np.ndarray:
    """1-D object array of strings encoded with the numcodecs VLenUTF8 filter: per chunk a uint32 item count followed
    by (uint32 length, utf-8 bytes) for each item."""
    shape = tuple(meta["shape"])
    chunks = tuple(meta["chunks"])
    if len(shape) != 1:
        raise ValueError("vlen-utf8 arrays are supported for 1-D string columns only")
    sep = meta.get("dimension_separator", ".")
    out = np.empty(shape, dtype=object)
    out[...] = ""
    for i in range((shape[0] + chunks[0] - 1) // chunks[0]):
        p = array_dir / str(i) if sep in (".", "/") else array_dir / str(i)
        if not p.exists():
            continue
        raw = _decode(p.read_bytes(), meta.get("compressor"), 0)
        n = int(np.frombuffer(raw, dtype="<u4", count=1)[0])
        off = 4
        for j in range(n):
            ln = int(np.frombuffer(raw, dtype="<u4", count=1, offset=off)[0])
            off += 4
            idx = i * chunks[0] + j
            if idx < shape[0]:
                out[idx] = raw[off:off + ln].decode("utf-8")
            off += ln
    return out


def read_array(array_dir: Path) -> np.ndarray:
    """Read a zarr-v2 array with any regular chunk grid (C or F order chunks, '.' or '/' chunk keys)."""
    array_dir = Path(array_dir)
    meta = json.loads((array_dir / ".zarray").read_text())
    filters = meta.get("filters") or []
    if [f.get("id") for f in filters] == ["vlen-utf8"]:
        return _read_vlen_utf8(array_dir, meta)       # how xarray stores string columns such as ANTENNA/NAME
    if filters:
        raise ValueError(f"{array_dir}: zarr filters {filters} are not supported")
    dtype = np.dtype(meta["dtype"])
    shape = tuple(meta["shape"])
    chunks = tuple(meta["chunks"]) if shape else ()
    codec = meta.get("compressor")
    order = meta.get("order", "C")
    sep = meta.get("dimension_separator", ".")
    fv = meta.get("fill_value")
    out = np.empty(shape, dtype=dtype)
    if fv is not None:
        try:
            if dtype.kind == "c" and isinstance(fv, list):
                out[...] = complex(float(fv[0]), float(fv[1]))
            elif dtype.kind in "fc":
                out[...] = float(fv)
            elif dtype.kind in "iu":
                out[...] = int(fv)
            else:
                out[...] = fv
        except (TypeError, ValueError):
            out[...] = 0
    else:
        out[...] = 0
    if not shape:
        p = array_dir / "0"
        if p.exists():
            out[...] = np.frombuffer(_decode(p.read_bytes(), codec, dtype.itemsize), dtype=dtype, count=1)[0]
        return out
    grid = [range((s + c - 1) // c) for s, c in zip(shape, chunks)]
    nbytes = int(np.prod(chunks)) * dtype.itemsize
    for idx in itertools.product(*grid):
        p = array_dir / sep.join(str(i) for i in idx)
        if not p.exists() and sep == ".":
            alt = array_dir.joinpath(*[str(i) for i in idx])
            p = alt if alt.exists() else p
        if not p.exists():
            continue  # missing chunk == fill_value
        raw = _decode(p.read_bytes(), codec, nbytes)
        block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunks))).reshape(chunks, order=order)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out


# ------------------------------------------------------------------------------------------------- leaves
def write_svd_to_zarr(svd_result, path, compressor: str, level: int, rowid: np.ndarray):
    """Same name, arguments and store layout as the reference's write_svd_to_zarr (compress_ms.py:723-763): a zarr-v2
    store rooted AT THE LEAF with U(time, mode), S(mode), WT(mode, channel) + coordinates time/mode/channel; the data
    variables use the requested compressor, the coordinates none."""
    U, s, V = svd_result
    # real factors (the rank-1 WEIGHT_SPECTRUM leaf, compress_ms.py:493-498) stay float32 as in the reference
    vdt = np.float32 if (np.isrealobj(np.asarray(U)) and np.isrealobj(np.asarray(V))) else np.complex64
    U = np.ascontiguousarray(np.asarray(U), dtype=vdt)
    s = np.ascontiguousarray(np.asarray(s), dtype=np.float32)
    V = np.ascontiguousarray(np.asarray(V), dtype=vdt)
    if U.ndim != 2 or V.ndim != 2 or s.ndim != 1 or U.shape[1] != s.shape[0] or V.shape[0] != s.shape[0]:
        raise ValueError(f"inconsistent factor shapes U{U.shape} S{s.shape} WT{V.shape}")
    rowid = np.asarray(rowid)
    if rowid.shape[0] != U.shape[0]:
        raise ValueError("rowid must have one entry per row of U")
    codec = get_compressor(compressor, level)
    arrays = {
        "U": (U, ("time", "mode")),
        "S": (s, ("mode",)),
        "WT": (V, ("mode", "channel")),
        "time": (rowid.astype(np.int64), ("time",)),
        "mode": (np.arange(s.shape[0], dtype=np.int64), ("mode",)),
        "channel": (np.arange(V.shape[1], dtype=np.int64), ("channel",)),
    }
    write_group(Path(path), arrays, attrs={}, codec_for=lambda n: codec if n in ("U", "S", "WT") else None)


def read_svd_from_zarr(path):
    """(U, S, WT, rowid) of one leaf — what the reference takes from xr.open_zarr(leaf) (decompress_ms.py:188-194).
    Works on leaves written by this module and on leaves written by the reference (xarray + zarr 2.18)."""
    path = Path(path)
    if not (path / "U" / ".zarray").exists():
        raise FileNotFoundError(f"{path} is not a factor leaf (no U array)")
    U = read_array(path / "U").astype(np.complex64, copy=False)
    S = read_array(path / "S").astype(np.float32, copy=False)
    WT = read_array(path / "WT").astype(np.complex64, copy=False)
    rowid = read_array(path / "time") if (path / "time" / ".zarray").exists() else np.arange(U.shape[0])
    if S.ndim == 2:
        S = S[:, 0]
    return U, S, WT, rowid


def list_subtables(zarr_path):
    """Directory names below a path (reference decompress_ms.py:76-92 list_subtables)."""
    zarr_path = str(zarr_path)
    if not os.path.isdir(zarr_path):
        return []
    return sorted(f for f in os.listdir(zarr_path) if os.path.isdir(os.path.join(zarr_path, f)))
