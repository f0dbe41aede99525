"""Dev probe (GPU): tridiagonalisation time of the MeerKAT shard for the tile-major copy ("tridiag_tiled") and L2 prefetch
distances ("tridiag_pf")."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B = 1040
A = torch.empty((B, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B // 4, 4)
ref = None
for variant in (1, 2):
    for tiled, pf in ((0, 0), (0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (1, 4)):
        eng.set_option("tridiag_variant", variant)
        eng.set_option("tridiag_pf", pf)
        eng.set_option("tridiag_tiled", tiled)
        eng.set_option("stage_timing", 1)
        ts = []
        for _ in range(3):
            res = eng.compress(A, decorrelation=0.99)
            torch.cuda.synchronize()
            ts.append(round(eng.last_eig_ms()["tridiag"], 2))
        eng.set_option("stage_timing", 0)
        S = res[1]
        if ref is None:
            ref = S.clone()
        dev = float(((S - ref).abs() / ref.abs().clamp_min(1e-20)).max())
        print(f"variant={variant} tiled={tiled} pf={pf}: tridiag ms {ts}  S identical to first: {bool(torch.equal(S, ref))} max rel dev {dev:.2e} ranks equal {bool(torch.equal(res[3], res[3]))}", flush=True)
