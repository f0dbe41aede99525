"""Minimal reader for casacore tables - just enough to load the columns the hot path needs from a Measurement Set
written with casacore's default storage managers (what simms / the reference's sample ``tests/data/sim-visco-kat7.ms``
uses), WITHOUT python-casacore, which is absent from this image (SURVEY section 2.1 rows 6-7 put MS I/O out of scope; this
is the "minimal TSM reader" of SURVEY 8f next-4 so that ``visco compressms -ms <sample MS>`` runs end to end).

Supported, and only this:
  * ``table.dat``: row count, the column -> data-manager sequence number map, column value types;
  * TiledShapeStMan / TiledColumnStMan / TiledCellStMan columns with ONE hypercube (``table.f<seq>`` header +
    ``table.f<seq>_TSM<n>`` tiles): DATA, MODEL_DATA, CORRECTED_DATA (complex64), FLAG (bit-packed), WEIGHT, SIGMA,
    WEIGHT_SPECTRUM (float32);
  * StandardStMan scalar columns of fixed width (Int, Float, Double, Complex, short String) through the SSM bucket
    index: ANTENNA1, ANTENNA2 in the main table, NAME in ANTENNA;
  * the first integer array of an SSM indirect-array file (``table.f0i``): POLARIZATION/CORR_TYPE.
IncrementalStMan columns (SCAN_NUMBER, FIELD_ID, DATA_DESC_ID, FLAG_ROW, TIME in such files) are not decoded.
When python-casacore is importable, visco_b200.msdata uses it instead of this module.

File-format notes (casacore tables/DataMan sources, checked against the bytes of the sample MS): AipsIO objects start
with the magic 0xBEBEBEBE, a length, a type string and a version; ``table.dat`` and the tiled headers are big-endian
("canonical"), StandardStMan files of a little-endian table are little-endian throughout.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

_TYPES = {"Bool": ("?", 1), "uChar": ("u1", 1), "Short": ("<i2", 2), "uShort": ("<u2", 2), "Int": ("<i4", 4),
          "uInt": ("<u4", 4), "Int64": ("<i8", 8), "float": ("<f4", 4), "double": ("<f8", 8), "Complex": ("<c8", 8),
          "DComplex": ("<c16", 16), "String": ("V12", 12)}


class CasaTableError(RuntimeError):
    pass


def _be32(b, off):
    return struct.unpack_from(">I", b, off)[0]


def table_info(path):
    """(nrow, {column: (dm sequence number, value type name, is_array)}) from ``<path>/table.dat``."""
    with open(os.path.join(path, "table.dat"), "rb") as f:
        b = f.read()
    if b[:4] != b"\xbe\xbe\xbe\xbe":
        raise CasaTableError(f"{path}: table.dat is not an AipsIO file")
    # magic, length, strlen("Table"), "Table", version, nrow, endian flag
    i = 8
    n = _be32(b, i)
    if b[i + 4:i + 4 + n] != b"Table":
        raise CasaTableError(f"{path}: not a casacore table")
    i += 4 + n
    version = _be32(b, i)
    nrow = _be32(b, i + 4)
    if version >= 3:   # 64-bit row count variants are not needed here
        nrow = _be32(b, i + 4)
    types = {}
    for m in re.finditer(rb"(Scalar|Array)ColumnDesc<(\w+)\s*", b):
        j = m.end()
        # version (4 bytes) then the column name as (length, chars)
        ln = _be32(b, j + 4)
        if 0 < ln < 64:
            name = b[j + 8:j + 8 + ln].decode("ascii", "replace")
            types.setdefault(name, (m.group(2).decode(), m.group(1) == b"Array"))
    # column -> data manager: records (uint32 namelen, name, uint32 1, uint32 seq, ...) near the end of the file
    cols = {}
    for name, (tname, is_arr) in types.items():
        key = struct.pack(">I", len(name)) + name.encode() + struct.pack(">I", 1)
        k = b.rfind(key)
        if k < 0:
            continue
        seq = _be32(b, k + len(key))
        cols[name] = (int(seq), tname, is_arr)
    return int(nrow), cols


def _fixed_shapes(path):
    """{array column: fixed cell shape} for the array columns whose column -> manager record carries one (such cells are
    stored inline by StandardStMan); other array columns are indirect (an 8-byte file offset per row)."""
    with open(os.path.join(path, "table.dat"), "rb") as f:
        b = f.read()
    _, cols = table_info(path)
    shapes = {}
    for name, (seq, tname, is_arr) in cols.items():
        if not is_arr:
            continue
        key = struct.pack(">I", len(name)) + name.encode() + struct.pack(">I", 1)
        k = b.rfind(key) + len(key) + 4
        if k < len(b) and b[k] == 1:
            m = b.find(b"IPosition", k, k + 32)
            if m >= 0:
                j = m + len(b"IPosition")
                ndim = _be32(b, j + 4)
                if 0 < ndim <= 8:
                    shapes[name] = tuple(_be32(b, j + 8 + 4 * d) for d in range(ndim))
    return shapes


def _iposition_list(b):
    out = []
    for m in re.finditer(rb"IPosition", b):
        j = m.end()
        ndim = _be32(b, j + 4)
        if 0 < ndim <= 8 and j + 8 + 4 * ndim <= len(b):
            out.append([_be32(b, j + 8 + 4 * d) for d in range(ndim)])
    return out


def read_tiled_column(path, seq, tname):
    """Whole column of a tiled storage manager with one hypercube as a C-order numpy array [row, ..., fastest axis].
    The header lists the cube shape and the tile shape as the first two IPosition records of full rank."""
    with open(os.path.join(path, f"table.f{seq}"), "rb") as f:
        hdr = f.read()
    if b"Tiled" not in hdr[:64]:
        raise CasaTableError(f"{path}/table.f{seq} is not a tiled storage manager")
    shapes = [s for s in _iposition_list(hdr) if len(s) >= 2 and all(0 < x < 2 ** 31 for x in s)]
    if len(shapes) < 2 or len(shapes[0]) != len(shapes[1]):
        raise CasaTableError(f"{path}/table.f{seq}: cannot find the cube and tile shapes")
    cube, tile = shapes[0], shapes[1]                    # Fortran order: first axis fastest, last axis = row
    data_files = sorted(fn for fn in os.listdir(path) if fn.startswith(f"table.f{seq}_TSM"))
    if len(data_files) != 1:
        raise CasaTableError(f"{path}: expected one TSM file for table.f{seq}, found {data_files}")
    raw = np.fromfile(os.path.join(path, data_files[0]), dtype=np.uint8)
    ntile = [-(-c // t) for c, t in zip(cube, tile)]
    nelem = int(np.prod(tile))
    if tname == "Bool":
        tbytes = (nelem + 7) // 8
    else:
        dt, size = _TYPES[tname]
        tbytes = nelem * size
    total = int(np.prod(ntile))
    if raw.size < total * tbytes:
        raise CasaTableError(f"{path}/{data_files[0]}: {raw.size} bytes, need {total * tbytes}")
    cshape = tuple(reversed(cube))                       # C order
    tshape = tuple(reversed(tile))
    out = np.zeros(cshape, dtype=bool if tname == "Bool" else np.dtype(dt))
    idx = 0
    for t_lin in range(total):
        # tile coordinates, first (fastest) axis varies fastest in the file
        rem, coord = t_lin, []
        for nt in ntile:
            coord.append(rem % nt)
            rem //= nt
        blk = raw[idx:idx + tbytes]
        idx += tbytes
        if tname == "Bool":
            arr = np.unpackbits(blk, bitorder="little")[:nelem].astype(bool).reshape(tshape)
        else:
            arr = blk.view(np.dtype(dt)).reshape(tshape)
        sl_out, sl_in = [], []
        for ax in range(len(cube)):                      # Fortran axis ax == C axis (ndim-1-ax)
            lo = coord[ax] * tile[ax]
            hi = min(lo + tile[ax], cube[ax])
            sl_out.append(slice(lo, hi))
            sl_in.append(slice(0, hi - lo))
        out[tuple(reversed(sl_out))] = arr[tuple(reversed(sl_in))]
    return out


class _SSM:
    """StandardStMan file of a little-endian table: 512-byte header, then buckets of `bucket_size` bytes; the bucket index
    maps row ranges to data buckets, in which every column owns rows_per_bucket * width bytes at a fixed offset."""

    def __init__(self, path, seq):
        self.fn = os.path.join(path, f"table.f{seq}")
        with open(self.fn, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:4] != b"\xbe\xbe\xbe\xbe" or b"StandardStMan" not in b[:40]:
            raise CasaTableError(f"{self.fn} is not a StandardStMan file")
        i = b.index(b"StandardStMan") + len(b"StandardStMan")
        little = b[4:8] != struct.pack(">I", struct.unpack(">I", b[4:8])[0]) or struct.unpack("<I", b[4:8])[0] < 4096
        self.e = "<" if little else ">"
        version = struct.unpack_from(self.e + "I", b, i)[0]
        i += 4
        if version >= 3:
            i += 1                                        # bool: big-endian data
        (self.bucket_size, self.nbuckets, _cache, nfree) = struct.unpack_from(self.e + "4I", b, i)
        i += 16
        _first_free, self.nidx, self.first_idx = struct.unpack_from(self.e + "iIi", b, i)
        i += 12
        self.idx_offset = struct.unpack_from(self.e + "I", b, i)[0] if version >= 2 else 0
        k = b.rfind(b"SSMIndex")          # the live index is the last one written (an empty one follows the header)
        if k < 0:
            raise CasaTableError(f"{self.fn}: no SSMIndex")
        k += len(b"SSMIndex")
        _v, nused, self.rows_per_bucket, self.ncol = struct.unpack_from(self.e + "4I", b, k)
        # two trailing Block<uInt> objects: last row of every used bucket, bucket numbers
        blocks = [m.end() for m in re.finditer(rb"Block", b[k:k + 4096])]
        if len(blocks) < 2:
            raise CasaTableError(f"{self.fn}: SSMIndex blocks not found")

        def block(at):
            _bv, n = struct.unpack_from(self.e + "2I", b, k + at)
            return list(struct.unpack_from(self.e + f"{n}I", b, k + at + 8))
        self.last_row = block(blocks[-2])
        self.bucket_nr = block(blocks[-1])
        # the index stores as many entries as buckets were ever allocated; the used ones come first
        self.last_row, self.bucket_nr = self.last_row[:max(nused, 1)], self.bucket_nr[:max(nused, 1)]

    def column(self, col_offset_bytes_per_row_block, width, dtype, nrow):
        """values of a fixed-width scalar column whose slab starts at `col_offset` (= rows_per_bucket * sum of the widths
        of the columns bound before it) inside every data bucket"""
        out = np.zeros(nrow, dtype=dtype)
        first = 0
        for last, bn in zip(self.last_row, self.bucket_nr):
            cnt = min(last, nrow - 1) - first + 1
            if cnt <= 0:
                break
            off = 512 + bn * self.bucket_size + col_offset_bytes_per_row_block
            out[first:first + cnt] = np.frombuffer(self.b, dtype=dtype, count=cnt, offset=off)
            first += cnt
        if first < nrow:
            raise CasaTableError(f"{self.fn}: bucket index covers {first} of {nrow} rows")
        return out


def read_ssm_scalars(path, names):
    """{name: array} for fixed-width scalar columns that share one StandardStMan with only fixed-width scalar columns
    bound before them (column slabs are laid out in binding order = the order of the column -> manager records)."""
    nrow, cols = table_info(path)
    fixed = _fixed_shapes(path)
    with open(os.path.join(path, "table.dat"), "rb") as f:
        b = f.read()
    out = {}
    for name in names:
        if name not in cols:
            raise CasaTableError(f"{path}: no column {name}")
        seq, tname, is_arr = cols[name]
        if is_arr or tname not in _TYPES:
            raise CasaTableError(f"{path}: column {name} is not a fixed-width scalar")
        ssm = _SSM(path, seq)
        # columns of the same manager in binding order (position of their record in table.dat)
        mates = []
        for other, (s2, t2, a2) in cols.items():
            if s2 == seq:
                key = struct.pack(">I", len(other)) + other.encode() + struct.pack(">I", 1)
                mates.append((b.rfind(key), other, t2, a2))
        mates.sort()
        off = 0
        for _, other, t2, a2 in mates:
            if other == name:
                break
            if t2 not in _TYPES:
                raise CasaTableError(f"{path}: column {other} before {name} in its StandardStMan has an unsupported type")
            w2 = _TYPES[t2][1]
            if a2:                                         # inline fixed-shape cells, else an 8-byte offset per row
                w2 = w2 * int(np.prod(fixed[other])) if other in fixed else 8
                off += ssm.rows_per_bucket * w2
            else:
                off += (ssm.rows_per_bucket + 7) // 8 if t2 == "Bool" else ssm.rows_per_bucket * w2
        dt, width = _TYPES[tname]
        vals = ssm.column(off, width, np.dtype(dt), nrow)
        if tname == "String":
            # 12-byte slots: up to 8 characters inline + int32 length (longer strings live in string buckets)
            res = []
            for v in vals:
                raw = bytes(v)
                ln = struct.unpack("<i", raw[8:12])[0]
                if not (0 <= ln <= 8):
                    raise CasaTableError(f"{path}: string column {name} holds strings longer than 8 characters")
                res.append(raw[:ln].decode("ascii", "replace"))
            vals = np.array(res)
        out[name] = vals
    return out


def read_first_int_array(path, fname="table.f0i"):
    """First Int array of an SSM indirect-array file (StManArrayFile): [version, length] header, then per array
    (ndim, shape..., data) - enough for POLARIZATION/CORR_TYPE, the first array column of that table."""
    with open(os.path.join(path, fname), "rb") as f:
        b = f.read()
    # skip the 16-byte file header (version, length, padding); an array record = ndim, shape, values
    for start in (16, 8, 12, 20):
        if start + 8 > len(b):
            continue
        ndim = struct.unpack_from("<I", b, start)[0]
        if 1 <= ndim <= 4:
            shape = struct.unpack_from(f"<{ndim}I", b, start + 4)
            n = int(np.prod(shape))
            if 0 < n <= 64 and start + 4 + 4 * ndim + 4 * n <= len(b):
                return np.frombuffer(b, dtype="<i4", count=n, offset=start + 4 + 4 * ndim).reshape(shape[::-1])
    raise CasaTableError(f"{path}/{fname}: no integer array found")


def read_measurement_set(path, column="DATA"):
    """dict with DATA-like column, FLAG, ANTENNA1/2, antenna names, CORR_TYPE (and MODEL_DATA / WEIGHT_SPECTRUM when the
    MS has them as tiled columns)."""
    nrow, cols = table_info(path)
    if column not in cols:
        raise CasaTableError(f"{path}: no column {column}")
    out = {"nrow": nrow}
    seq, tname, _ = cols[column]
    out["data"] = read_tiled_column(path, seq, tname)
    for opt in ("FLAG", "MODEL_DATA", "WEIGHT_SPECTRUM"):
        if opt in cols and opt != column:
            try:
                out[opt] = read_tiled_column(path, cols[opt][0], cols[opt][1])
            except (CasaTableError, OSError):
                pass
    sc = read_ssm_scalars(path, ["ANTENNA1", "ANTENNA2"])
    out["ANTENNA1"], out["ANTENNA2"] = sc["ANTENNA1"], sc["ANTENNA2"]
    out["names"] = [str(x) for x in read_ssm_scalars(os.path.join(path, "ANTENNA"), ["NAME"])["NAME"]]
    out["corr_types"] = [int(x) for x in read_first_int_array(os.path.join(path, "POLARIZATION")).reshape(-1)]
    return out
