"""Dev probe (GPU): configs[3] cube (8320 x 64 x 64) compress time and eigensolver kernel times for the row-group count of
tridiag_small_kernel ("tridiag_small_rs" = 1, 2, 4); singular values against rs = 1."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
for (B, m, n, kw) in [(8320, 64, 64, dict(compressionrank=8)), (2000, 48, 80, dict(decorrelation=0.95)), (2000, 90, 57, dict(compressionrank=6))]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B // 4, 4)
    ref = None
    for rs in (1, 2, 4):
        eng.set_option("tridiag_small_rs", rs)
        res = eng.compress(A, **kw)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            res = eng.compress(A, **kw)
        e1.record(); e1.synchronize()
        t = e0.elapsed_time(e1) / 5
        eng.set_option("stage_timing", 1)
        eng.compress(A, **kw)
        torch.cuda.synchronize()
        em = eng.last_eig_ms()
        eng.set_option("stage_timing", 0)
        S, rk = res[1].clone(), res[3].clone()
        if ref is None:
            ref = (S, rk)
        same = rk == ref[1]
        dev = float((((S - ref[0]).abs() / ref[0].abs().clamp_min(1e-20)).amax(dim=1))[same].max())
        print(f"{B} x {m} x {n} {kw} rs={rs}: compress {t:.3f} ms, tridiag {em['tridiag']:.3f} ms, rank mismatches {int((~same).sum())}, max rel dS vs rs=1 {dev:.2e}", flush=True)
eng.set_option("tridiag_small_rs", 0)
