"""Build tests/golden/sample_ms_kat7.npz from the reference's sample Measurement Set (run here; needs /root/reference).

Decodes the DATA column of tests/data/sim-visco-kat7.ms without casacore (TiledShapeStMan file table.f2_TSM1: 76 tiles
of [100 rows][16 chan][4 corr] little-endian complex64, 7560 valid rows = 360 timeslots x 21 baselines, time-major,
baselines (0,1),(0,2)...(5,6); antennas ANT-0..ANT-6) and keeps the rows of the 6 baselines among antennas 0..3 to stay
small (360 x 6 rows). The bundle format is visco_b200.msdata.VisData's."""
import os
import sys
from itertools import combinations

import numpy as np

REF = os.environ.get("VISCO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from visco_b200.msdata import VisData  # noqa: E402

raw = np.fromfile(os.path.join(REF, "tests/data/sim-visco-kat7.ms/table.f2_TSM1"), dtype="<c8")
cube = raw.reshape(76 * 100, 16, 4)[:7560]
pairs = list(combinations(range(7), 2))
a1 = np.tile(np.array([p[0] for p in pairs], np.int32), 360)
a2 = np.tile(np.array([p[1] for p in pairs], np.int32), 360)
keep = (a1 < 4) & (a2 < 4)
rowid = np.nonzero(keep)[0].astype(np.int64)
vis = VisData(data=cube[keep], antenna1=a1[keep], antenna2=a2[keep], antenna_names=[f"ANT-{i}" for i in range(7)],
              corr_types=[9, 10, 11, 12], rowid=rowid)
out = os.path.join(HERE, "sample_ms_kat7.npz")
vis.save(out)
print(out, os.path.getsize(out), vis.data.shape, len(vis.baselines()))
