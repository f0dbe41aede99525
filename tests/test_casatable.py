"""CPU: the minimal casacore-table reader (visco_b200.casatable) against the subset of the reference's sample Measurement
Set committed under tests/golden (tests/golden/make_sample_ms_subset.py) and, when /root/reference is present, against
the original; plus the driver-level checks the reference's own test configuration relies on."""
import os
from itertools import combinations

import numpy as np
import pytest

from visco_b200 import casatable as ct
from visco_b200.msdata import VisData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBSET = os.path.join(ROOT, "tests", "golden", "sim-visco-kat7-subset.ms")
REFMS = "/root/reference/tests/data/sim-visco-kat7.ms"


@pytest.mark.parametrize("path", [SUBSET, REFMS])
def test_sample_ms_columns(path):
    if not os.path.isdir(path):
        pytest.skip(f"{path} not present")
    nrow, cols = ct.table_info(path)
    assert nrow == 7560 and cols["DATA"][1:] == ("Complex", True) and cols["ANTENNA1"][1:] == ("Int", False)
    ms = ct.read_measurement_set(path)
    assert ms["data"].shape == (7560, 16, 4) and ms["data"].dtype == np.complex64
    assert ms["FLAG"].shape == (7560, 16, 4) and ms["FLAG"].dtype == bool and not ms["FLAG"].any()
    pairs = list(combinations(range(7), 2))                       # 360 timeslots x 21 baselines, time-major
    np.testing.assert_array_equal(ms["ANTENNA1"], np.tile([a for a, _ in pairs], 360))
    np.testing.assert_array_equal(ms["ANTENNA2"], np.tile([b for _, b in pairs], 360))
    assert ms["names"] == [f"ANT-{i}" for i in range(7)] and ms["corr_types"] == [9, 10, 11, 12]
    # the tiles are [100 rows][16 chan][4 corr] little-endian complex64, rows beyond 7560 are padding
    raw = np.fromfile(os.path.join(path, "table.f2_TSM1"), dtype="<c8").reshape(76 * 100, 16, 4)[:7560]
    np.testing.assert_array_equal(ms["data"], raw)
    # the bundle committed in round 1 was decoded from the same file
    with np.load(os.path.join(ROOT, "tests", "golden", "sample_ms_kat7.npz")) as z:
        np.testing.assert_array_equal(ms["data"][z["ROWID"]], z["DATA"])


def test_visdata_loads_a_measurement_set_without_casacore():
    vis = VisData.load(SUBSET, column="DATA", scan=1, fieldid=0, ddid=0)
    assert vis.data.shape == (7560, 16, 4) and len(vis.baselines()) == 21
    assert vis.baseline_rows(0, 1).size == 360 and vis.corr_index("YY") == 3
    with pytest.raises(RuntimeError):
        VisData.load(os.path.join(ROOT, "tests"), column="DATA")          # a directory that is not a table
    with pytest.raises(RuntimeError):
        VisData.load(SUBSET, column="NO_SUCH_COLUMN")


def test_generic_tile_walk():
    """tiles that cut more than the row axis, with partial edge tiles, bit-packed booleans"""
    import struct
    import tempfile
    cube, tile = (3, 5, 11), (2, 4, 4)                                    # Fortran order (corr, chan, row)
    full = np.arange(np.prod(cube), dtype=np.float32).reshape(tuple(reversed(cube)))
    ntile = [-(-c // t) for c, t in zip(cube, tile)]
    with tempfile.TemporaryDirectory() as d:
        def ipos(v):
            return b"IPosition" + struct.pack(">I", 1) + struct.pack(">I", len(v)) + b"".join(struct.pack(">I", x) for x in v)
        open(os.path.join(d, "table.f4"), "wb").write(b"\xbe\xbe\xbe\xbe" + b"\0" * 8 + b"TiledShapeStMan" + ipos(cube) + ipos(tile))
        blobs, bits = [], []
        for t2 in range(ntile[2]):
            for t1 in range(ntile[1]):
                for t0 in range(ntile[0]):
                    blk = np.zeros(tuple(reversed(tile)), np.float32)
                    src = full[t2 * 4:(t2 + 1) * 4, t1 * 4:(t1 + 1) * 4, t0 * 2:(t0 + 1) * 2]
                    blk[:src.shape[0], :src.shape[1], :src.shape[2]] = src
                    blobs.append(blk.tobytes())
                    bits.append(np.packbits((blk.reshape(-1) % 3 == 0), bitorder="little").tobytes())
        open(os.path.join(d, "table.f4_TSM0"), "wb").write(b"".join(blobs))
        np.testing.assert_array_equal(ct.read_tiled_column(d, 4, "float"), full)
        os.remove(os.path.join(d, "table.f4_TSM0"))
        open(os.path.join(d, "table.f4_TSM0"), "wb").write(b"".join(bits))
        np.testing.assert_array_equal(ct.read_tiled_column(d, 4, "Bool"), full % 3 == 0)
