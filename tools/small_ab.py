"""Dev probe (GPU): BASELINE configs[3] (8320 matrices of 64 x 64, fixed rank) through the one-sided Jacobi path and through the Gram path."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B, m, n, k = 8320, 64, 64, 8
A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B // 4, 4)
ref = None
for off in (0, 1, 1, 0):                   # 0: one-sided Jacobi route (small_impl = 1), 1: default route
    eng.set_option("small_impl", 0 if off else 1)
    eng.set_option("stage_timing", 2 if off else 0)
    for _ in range(2):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    if ref is None:
        ref = S.clone()
    dev = float(((S - ref).abs() / ref.abs().clamp_min(1e-20)).max())
    rec = eng.reconstruct(U, S, Vt, ranks)
    err = float((A - rec).abs().pow(2).sum().sqrt() / A.abs().pow(2).sum().sqrt())
    print(f"small_off={off}: compress {ms:.3f} ms  max rel dS {dev:.2e}  rel recon err {err:.6f}", flush=True)
