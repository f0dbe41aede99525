"""Dev probe (GPU): per-kernel event times of the direct eigensolver (stage_timing=2 prints them on stderr;
stage_timing=1 runs the scalar QL kernel on the main stream instead of the side stream)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
a = [int(x) for x in sys.argv[1:]] or [112, 256, 1024]
for i in range(0, len(a), 3):
    B, m, n = a[i:i + 3]
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B, 1)
    G = eng.gram(A)
    for impl, mode in ((2, 0), (2, 2), (2, 0), (2, 0), (2, 1), (2, 1), (1, 0), (1, 0)):
        eng.set_option("eig_impl", impl)
        eng.set_option("stage_timing", mode)
        W = G.clone()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.eigh_jacobi(W)
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B} r={m} eig_impl={impl} stage_timing={mode}: eigh total {e0.elapsed_time(e1):.3f} ms", flush=True)
    eng.set_option("stage_timing", 0)
    eng.set_option("eig_impl", 1)
