"""Dev: precision of the Jacobi rotations — full-rank factorisation residual and factor orthonormality."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
for key, val in [a.split("=") for a in sys.argv[1:]]:
    eng.set_option(key, float(val))
for (B, m, n) in [(8, 64, 64), (8, 360, 16), (4, 160, 400), (4, 256, 1024)]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B, 1, nbl_total=8)
    U, S, Vt, ranks, stats = eng.compress(A)
    torch.cuda.synchronize()
    a = A.cpu().numpy().astype(np.complex128)
    u, s, vt = U.cpu().numpy().astype(np.complex128), S.cpu().numpy().astype(np.float64), Vt.cpu().numpy().astype(np.complex128)
    res = max(np.linalg.norm(a[b] - (u[b] * s[b]) @ vt[b]) / np.linalg.norm(a[b]) for b in range(B))
    r = min(m, n)
    ou = max(np.abs(u[b].conj().T @ u[b] - np.eye(r)).max() for b in range(B))
    ov = max(np.abs(vt[b] @ vt[b].conj().T - np.eye(r)).max() for b in range(B))
    sref = np.stack([np.linalg.svd(a[b], compute_uv=False) for b in range(B)])
    serr = np.max(np.abs(s - sref) / sref)
    print(f"{m}x{n}: residual {res:.2e}  U orth {ou:.2e}  Vt orth {ov:.2e}  sigma rel {serr:.2e}  sweeps {float(stats[:,2].mean()):.1f}", flush=True)
