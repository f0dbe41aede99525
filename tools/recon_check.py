"""Dev check/timing of the reconstruction kernels (run on the GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
HBM = 6548.2


def bench(B, m, n, k, generic):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    U = torch.view_as_complex(torch.randn((B, m, k, 2), device=dev, generator=g))
    Vt = torch.view_as_complex(torch.randn((B, k, n, 2), device=dev, generator=g))
    S = torch.rand((B, k), device=dev, generator=g) + 0.5
    out = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
    eng.set_option("recon_generic", generic)
    for _ in range(3):
        eng.reconstruct(U, S, Vt, None, out=out)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    t0.record()
    for _ in range(reps):
        eng.reconstruct(U, S, Vt, None, out=out)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    ref = (U[:2] * S[:2, None, :]) @ Vt[:2]
    err = float((out[:2] - ref).abs().max() / ref.abs().max())
    by = B * (8.0 * m * n + 8.0 * k * (m + n) + 4.0 * k)
    print(f"B={B} m={m} n={n} k={k} generic={generic}: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / HBM:.2f} of HBM)  err {err:.1e}", flush=True)


for (B, m, n) in [(112, 256, 1024), (1024, 128, 2048), (1040, 512, 4096) if False else (256, 512, 4096)]:
    for k in (1, 2, 4, 8):
        bench(B, m, n, k, 0)
    bench(B, m, n, 8, 1)
    bench(B, m, n, 16, 0)
