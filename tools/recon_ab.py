"""Development tool: reconstruction time at the BASELINE configs[4] shape for the persistent tcgen05 kernel vs the older ones."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
B, m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 128, 2048
out = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
for k in (9, 10, 12, 14, 16, 17, 20, 24, 28, 31, 32):
    U = torch.randn((B, m, k), dtype=torch.complex64, device="cuda:0") / (2 * m) ** 0.5
    Vt = torch.randn((B, k, n), dtype=torch.complex64, device="cuda:0") / (2 * n) ** 0.5
    S = torch.rand((B, k), dtype=torch.float32, device="cuda:0") + 0.5
    ref = None
    for impl in (1, 0):
        eng.set_option("recon_tc_impl", impl)
        for _ in range(2):
            eng.reconstruct(U, S, Vt, None, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.reconstruct(U, S, Vt, None, out=out)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        by = B * (8.0 * m * n + 8.0 * k * (m + n) + 4 * k)
        chk = out[:4].clone()
        if ref is None:
            ref = chk
            err = 0.0
        else:
            err = float((chk - ref).abs().max() / ref.abs().max())
        print(f"k={k:2d} impl={impl}: {ms:7.3f} ms  {by / ms / 1e6:7.0f} GB/s  frac {by / ms / 1e6 / 6548.2:.3f}  maxdiff vs other impl {err:.2e}", flush=True)
    eng.set_option("recon_tc_impl", 0)
