"""Dev probe (GPU): error of the smallest retained singular value through the float32 Gram path as a function of its
ratio to the largest one (option illcond_thr = 0 disables the Gram-free redo)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
eng.set_option("illcond_thr", 0.0)
rng = np.random.default_rng(0)
for (m, n) in [(128, 256), (256, 1024)]:
    Q1, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    Q2, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    for rho in [0.05, 0.03, 0.02, 0.01, 0.007, 0.005, 0.003, 0.002, 0.001, 0.0005]:
        worst = 0.0
        for trial in range(4):
            sv = np.concatenate([1.0 + rng.random(4), rho * (0.5 + rng.random(m - 4))])
            sv = np.sort(sv)[::-1]
            A = ((Q1 * sv[None, :]) @ Q2[:m]).astype(np.complex64)
            ref = np.linalg.svd(A.astype(np.complex128), compute_uv=False)
            k = 24
            U, S, Vt, ranks, stats = eng.compress(torch.from_numpy(A[None]).cuda(), compressionrank=k)
            S = S[0].cpu().numpy()
            worst = max(worst, float(np.max(np.abs(S - ref[:k]) / ref[:k])))
        tol = 1e-4 + 2e-6 / (sv[k - 1] / sv[0])
        print(f"{m}x{n} rho={rho:7.4f}  sigma_k/sigma_1={sv[k-1]/sv[0]:.4f}  worst rel err {worst:.2e}  (parity tolerance {tol:.2e})", flush=True)
