"""Dev probe (GPU): tridiagonalisation time of the MeerKAT shard for L2 prefetch distances ("tridiag_pf") under both launch
shapes ("tridiag_variant" 1 = one matrix per SM, 2 = two)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B = 1040
A = torch.empty((B, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B // 4, 4)
for variant in (1, 2):
    for pf in (0, 1, 2, 4, 8):
        eng.set_option("tridiag_variant", variant)
        eng.set_option("tridiag_pf", pf)
        eng.set_option("stage_timing", 1)
        ts = []
        for _ in range(3):
            eng.compress(A, decorrelation=0.99)
            torch.cuda.synchronize()
            ts.append(round(eng.last_eig_ms()["tridiag"], 2))
        eng.set_option("stage_timing", 0)
        print(f"variant={variant} pf={pf}: tridiag ms {ts}", flush=True)
eng.set_option("tridiag_variant", 0)
eng.set_option("tridiag_pf", 1)
