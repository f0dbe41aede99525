"""Minimal in-memory view of the Measurement-Set columns the hot path touches, and a portable bundle format for it.

The reference reads a casacore Measurement Set through dask-ms (visco/compress_ms.py:54-194) and writes one back
(visco/decompress_ms.py:329-402); both libraries are outside the scope of this build and
absent from the image. What the hot path needs from the MAIN table is small: the visibility column
``DATA[row, chan, corr]``, ``ANTENNA1/2[row]``, ``ROWID``, the antenna names and the correlation types. ``VisData``
holds exactly that; it loads from / saves to an ``.npz`` bundle, or — when python-casacore/dask-ms are importable — from
a real MS. The callers on either side of the SVD (baseline gather, leaf tree, scatter) work on ``VisData``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# casacore Stokes enum <-> name (reference visco/ms_corr_types.yaml:1-34)
CORR_TYPES = {"Undefined": 0, "I": 1, "Q": 2, "U": 3, "V": 4, "RR": 5, "RL": 6, "LR": 7, "LL": 8, "XX": 9, "XY": 10,
              "YX": 11, "YY": 12, "RX": 13, "RY": 14, "LX": 15, "LY": 16, "XR": 17, "XL": 18, "YR": 19, "YL": 20,
              "PP": 21, "PQ": 22, "QP": 23, "QQ": 24, "RCircular": 25, "LCircular": 26, "Linear": 27, "Ptotal": 28,
              "Plinear": 29, "PFtotal": 30, "PFlinear": 31, "Pangle": 32}
CORR_TYPES_REVERSE = {v: k for k, v in CORR_TYPES.items()}


@dataclass
class VisData:
    data: np.ndarray                 # [row, chan, corr] complex64
    antenna1: np.ndarray             # [row] int32
    antenna2: np.ndarray             # [row] int32
    antenna_names: list              # [nant] str
    corr_types: list = field(default_factory=lambda: [9, 10, 11, 12])   # casacore enums of the corr axis
    rowid: np.ndarray | None = None  # [row] int64
    flag: np.ndarray | None = None   # [row, chan, corr] bool (optional)
    flag_row: np.ndarray | None = None       # [row] bool (optional)
    model_data: np.ndarray | None = None     # [row, chan, corr] complex64 (optional; flag replacement source)
    weight_spectrum: np.ndarray | None = None  # [row, chan, corr] float32 (optional; compressed to rank 1)
    sigma_spectrum: np.ndarray | None = None   # only set by the decompressor (reference decompress_ms.py:263-269)
    column: str = "DATA"

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data, dtype=np.complex64)
        if self.data.ndim != 3:
            raise ValueError("DATA must be [row, chan, corr]")
        nrow = self.data.shape[0]
        self.antenna1 = np.asarray(self.antenna1, dtype=np.int32)
        self.antenna2 = np.asarray(self.antenna2, dtype=np.int32)
        if self.antenna1.shape != (nrow,) or self.antenna2.shape != (nrow,):
            raise ValueError("ANTENNA1/ANTENNA2 must have one entry per row")
        if self.rowid is None:
            self.rowid = np.arange(nrow, dtype=np.int64)
        self.rowid = np.asarray(self.rowid, dtype=np.int64)
        self.antenna_names = [str(x) for x in self.antenna_names]
        self.corr_types = [int(c) for c in self.corr_types]
        if len(self.corr_types) != self.data.shape[2]:
            raise ValueError("corr_types must describe the correlation axis")
        if self.flag is not None:
            self.flag = np.ascontiguousarray(self.flag, dtype=bool)
            if self.flag.shape != self.data.shape:
                raise ValueError("FLAG must have the shape of the visibility column")
        if self.flag_row is not None:
            self.flag_row = np.ascontiguousarray(self.flag_row, dtype=bool)
        if self.model_data is not None:
            self.model_data = np.ascontiguousarray(self.model_data, dtype=np.complex64)
            if self.model_data.shape != self.data.shape:
                raise ValueError("MODEL_DATA must have the shape of the visibility column")
        if self.weight_spectrum is not None:
            self.weight_spectrum = np.ascontiguousarray(self.weight_spectrum, dtype=np.float32)
            if self.weight_spectrum.ndim != 3 or self.weight_spectrum.shape[0] != nrow:
                raise ValueError("WEIGHT_SPECTRUM must be [row, chan, corr]")

    # --------------------------------------------------------------------------------------------- persistence
    def save(self, path: str):
        np.savez_compressed(path, DATA=self.data, ANTENNA1=self.antenna1, ANTENNA2=self.antenna2, ROWID=self.rowid,
                            ANTENNA_NAME=np.array(self.antenna_names), CORR_TYPE=np.array(self.corr_types, np.int32),
                            **({"FLAG": self.flag} if self.flag is not None else {}),
                            **({"FLAG_ROW": self.flag_row} if self.flag_row is not None else {}),
                            **({"MODEL_DATA": self.model_data} if self.model_data is not None else {}),
                            **({"WEIGHT_SPECTRUM": self.weight_spectrum} if self.weight_spectrum is not None else {}),
                            **({"SIGMA_SPECTRUM": self.sigma_spectrum} if self.sigma_spectrum is not None else {}))

    @classmethod
    def load(cls, path: str, column: str = "DATA", scan=None, fieldid=None, ddid=None):
        """``.npz`` bundle, or a Measurement Set when dask-ms / python-casacore are available."""
        if str(path).endswith(".npz"):
            if not os.path.exists(path):
                raise ValueError(f"Measurement Set bundle {path} does not exist")
            with np.load(path, allow_pickle=False) as z:
                key = column if column in z.files else "DATA"
                keep = None
                for col, val in (("SCAN_NUMBER", scan), ("FIELD_ID", fieldid), ("DATA_DESC_ID", ddid)):
                    if col in z.files and val is not None:
                        ok = z[col] == int(val)
                        if not ok.any():
                            raise ValueError(f"Invalid selected {col} {val}. Available: {np.unique(z[col]).tolist()}")   # reference :461-468
                        keep = ok if keep is None else (keep & ok)
                if keep is not None and not keep.all():
                    def rows(name):
                        return z[name][keep] if name in z.files else None
                    return cls(data=z[key][keep], antenna1=z["ANTENNA1"][keep], antenna2=z["ANTENNA2"][keep],
                               antenna_names=list(z["ANTENNA_NAME"]), corr_types=list(z["CORR_TYPE"]), rowid=z["ROWID"][keep],
                               flag=rows("FLAG"), flag_row=rows("FLAG_ROW"), model_data=rows("MODEL_DATA"),
                               weight_spectrum=rows("WEIGHT_SPECTRUM"), column=column)
                return cls(data=z[key], antenna1=z["ANTENNA1"], antenna2=z["ANTENNA2"],
                           antenna_names=list(z["ANTENNA_NAME"]), corr_types=list(z["CORR_TYPE"]), rowid=z["ROWID"],
                           flag=z["FLAG"] if "FLAG" in z.files else None,
                           flag_row=z["FLAG_ROW"] if "FLAG_ROW" in z.files else None,
                           model_data=z["MODEL_DATA"] if "MODEL_DATA" in z.files else None,
                           weight_spectrum=z["WEIGHT_SPECTRUM"] if "WEIGHT_SPECTRUM" in z.files else None, column=column)
        if not os.path.exists(path):
            raise ValueError(f"Measurement Set {path} does not exist")     # reference compress_ms.py:876-877
        try:
            from casacore.tables import table  # type: ignore
        except ImportError:
            return cls._load_without_casacore(path, column, scan, fieldid, ddid)
        t = table(path, ack=False)
        q = []
        if scan is not None:
            q.append(f"SCAN_NUMBER=={int(scan)}")
        if fieldid is not None:
            q.append(f"FIELD_ID=={int(fieldid)}")
        if ddid is not None:
            q.append(f"DATA_DESC_ID=={int(ddid)}")
        if q:
            t = t.query(" && ".join(q))
            if t.nrows() == 0:
                raise ValueError("Invalid selection: no rows match scan/field/ddid")    # reference :461-468
        names = list(table(os.path.join(path, "ANTENNA"), ack=False).getcol("NAME"))
        corr = list(table(os.path.join(path, "POLARIZATION"), ack=False).getcol("CORR_TYPE")[0])
        have = set(t.colnames())
        # ROWID as dask-ms defines it: the row number in the parent table (what the leaves store as `time`)
        return cls(data=t.getcol(column), antenna1=t.getcol("ANTENNA1"), antenna2=t.getcol("ANTENNA2"), antenna_names=names,
                   corr_types=corr, rowid=np.asarray(t.rownumbers(), dtype=np.int64), column=column,
                   flag=t.getcol("FLAG") if "FLAG" in have else None,
                   flag_row=t.getcol("FLAG_ROW") if "FLAG_ROW" in have else None,
                   model_data=t.getcol("MODEL_DATA") if ("MODEL_DATA" in have and column != "MODEL_DATA") else None,
                   weight_spectrum=t.getcol("WEIGHT_SPECTRUM") if "WEIGHT_SPECTRUM" in have else None)

    @classmethod
    def _load_without_casacore(cls, path, column, scan, fieldid, ddid):
        """Measurement Set through visco_b200.casatable (tiled + StandardStMan columns only). The selection columns
        (SCAN_NUMBER / FIELD_ID / DATA_DESC_ID) and FLAG_ROW sit in an IncrementalStMan in such files and are not
        decoded: the whole table is taken, with a warning when a selection was asked for."""
        from . import LOG
        from .casatable import CasaTableError, read_measurement_set
        try:
            ms = read_measurement_set(path, column)
        except (CasaTableError, OSError, KeyError, ValueError) as e:
            raise RuntimeError(
                f"cannot read {path} without python-casacore ({e}); convert the MS columns to an .npz bundle "
                "(visco_b200.msdata.VisData.save) or install python-casacore") from e
        if any(v not in (None, 0, 1) for v in (scan, fieldid, ddid)):
            LOG.warning("python-casacore is not installed: SCAN_NUMBER / FIELD_ID / DATA_DESC_ID cannot be read, the whole "
                        "table is compressed as one scan / field / data description")
        return cls(data=ms["data"], antenna1=ms["ANTENNA1"], antenna2=ms["ANTENNA2"], antenna_names=ms["names"],
                   corr_types=ms["corr_types"], rowid=np.arange(ms["nrow"], dtype=np.int64), column=column,
                   flag=ms.get("FLAG"), model_data=ms.get("MODEL_DATA") if column != "MODEL_DATA" else None,
                   weight_spectrum=ms.get("WEIGHT_SPECTRUM"))

    # --------------------------------------------------------------------------------------------- hot-path helpers
    def baselines(self, antennas=None):
        """Unique (min, max) antenna pairs, autocorrelations excluded (reference compress_ms.py:508-520)."""
        if antennas:
            from itertools import combinations
            return list(combinations([int(a) for a in antennas], 2))
        a = np.minimum(self.antenna1, self.antenna2).astype(np.int64)
        b = np.maximum(self.antenna1, self.antenna2).astype(np.int64)
        keep = a != b
        pairs = np.unique(np.stack([a[keep], b[keep]], axis=1), axis=0)
        return [(int(p[0]), int(p[1])) for p in pairs]

    def baseline_rows(self, a1: int, a2: int) -> np.ndarray:
        """Row indices with ANTENNA1 == a1 and ANTENNA2 == a2 (reference compress_ms.py:591, decompress_ms.py:179-180)."""
        return np.nonzero((self.antenna1 == a1) & (self.antenna2 == a2))[0]

    def corr_index(self, name_or_enum) -> int:
        enum = CORR_TYPES[str(name_or_enum)] if not isinstance(name_or_enum, (int, np.integer)) else int(name_or_enum)
        try:
            return self.corr_types.index(enum)
        except ValueError:
            raise ValueError(f"correlation {name_or_enum} is not in this data set ({self.corr_types})") from None
