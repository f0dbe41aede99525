"""Dev check of the tcgen05 Gram kernel against an fp64 Gram (run under `timeout` on the GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
ok = True
for (B, m, n) in [(2, 128, 256), (3, 256, 1024), (2, 200, 300), (2, 512, 1024), (2, 130, 130), (1, 96, 2050), (5, 384, 640)]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B, 1, nbl_total=8)
    torch.cuda.synchronize()
    W = eng.gram(A, impl=2)
    torch.cuda.synchronize()
    Ws = eng.gram(A, impl=1)
    torch.cuda.synchronize()
    a = A.cpu().numpy().astype(np.complex128)
    G = np.einsum("btv,biv->bit", a, a.conj())
    e_tc = np.abs(W.cpu().numpy() - G).max() / np.abs(G).max()
    e_simt = np.abs(Ws.cpu().numpy() - G).max() / np.abs(G).max()
    print(f"B={B} m={m} n={n}: tcgen05 err {e_tc:.3e}  simt err {e_simt:.3e}", flush=True)
    ok &= e_tc < 2e-6
# timing at the C2 / C3 shapes
for (B, m, n) in [(112, 256, 1024), (32, 512, 4096)]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B, 1, nbl_total=64)
    for impl in (2, 1):
        for _ in range(2):
            eng.gram(A, impl=impl)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            eng.gram(A, impl=impl)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        fl = 8.0 * m * m * n * B
        print(f"gram impl={impl} B={B} m={m} n={n}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)
print("OK" if ok else "FAIL")
