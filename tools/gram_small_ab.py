"""Dev probe (GPU): fused small Gram + normalisation kernel ("gram_small" = 0) against the SIMT GEMM + normalisation pass (1):
singular values of both routes on several shapes with min(m, n) <= 64, and the compress time of the configs[3] cube."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
for (B, m, n, kw) in [(8320, 64, 64, dict(compressionrank=8)), (1000, 48, 100, dict(decorrelation=0.95)), (500, 100, 48, dict(compressionrank=5)),
                      (300, 64, 200, dict(decorrelation=0.9)), (200, 33, 64, dict(compressionrank=4)), (64, 130, 40, dict(compressionrank=40)),
                      (8320, 64, 64, dict(decorrelation=0.95))]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B // 4, 4)
    out = {}
    for impl in (1, 0):
        eng.set_option("gram_small", impl)
        res = eng.compress(A, **kw)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            res = eng.compress(A, **kw)
        e1.record(); e1.synchronize()
        out[impl] = (res[1].clone(), res[3].clone(), e0.elapsed_time(e1) / 5)
    S1, r1, t1 = out[1]
    S0, r0, t0 = out[0]
    same_rank = r1 == r0
    dev = float((((S0 - S1).abs() / S1.abs().clamp_min(1e-20)).amax(dim=1))[same_rank].max())
    print(f"{B} x {m} x {n} {kw}: old {t1:.3f} ms, fused {t0:.3f} ms, rank mismatches {int((~same_rank).sum())}, max rel dS {dev:.2e}", flush=True)
eng.set_option("gram_small", 0)
