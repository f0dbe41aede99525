"""Copy the table files visco_b200.casatable reads (and nothing else) from the reference's sample Measurement Set into
tests/golden/sim-visco-kat7-subset.ms, so that the `visco compressms -ms <MS>` end-to-end test can run on the GPU box, where
/root/reference does not exist. Run here (needs /root/reference). Data fixture only: main-table description, the
StandardStMan file with ANTENNA1 / ANTENNA2, the tiled DATA and FLAG columns, ANTENNA and POLARIZATION subtables.
MODEL_DATA / CORRECTED_DATA (all zeros in this MS, 3.9 MB each) and the other subtables are left out."""
import os
import shutil

REF = os.environ.get("VISCO_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "tests/data/sim-visco-kat7.ms")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim-visco-kat7-subset.ms")
FILES = ["table.dat", "table.info", "table.f1", "table.f2", "table.f2_TSM1", "table.f3", "table.f3_TSM1",
         "ANTENNA/table.dat", "ANTENNA/table.info", "ANTENNA/table.f0",
         "POLARIZATION/table.dat", "POLARIZATION/table.info", "POLARIZATION/table.f0", "POLARIZATION/table.f0i"]
for f in FILES:
    os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
print(DST, sum(os.path.getsize(os.path.join(DST, f)) for f in FILES), "bytes")
