"""Dev probe (GPU): graded ('signal') matrices at large sizes, full rank, default settings vs forced paths."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
rng = np.random.default_rng(11)
g = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
for (m, n) in [(300, 320), (586, 627), (520, 900)]:
    for trial in range(3):
        k = int(rng.integers(1, 100))
        a = (g(m, k) @ (np.diag(10.0 ** rng.uniform(0, 2, k)) @ g(k, n)) / np.sqrt(k) + 0.1 * g(m, n)).astype(np.complex64)
        ref = np.linalg.svd(a.astype(np.complex128), compute_uv=False)
        Ad = torch.from_numpy(a[None]).cuda()
        for thr in (0.005, 0.0, 1.0):
            eng.set_option("illcond_thr", thr)
            U, S, Vt, ranks, stats = eng.compress(Ad)
            torch.cuda.synchronize()
            Uh, Sh, Vh, st = U[0].cpu().numpy(), S[0].cpu().numpy(), Vt[0].cpu().numpy(), stats[0].cpu().numpy()
            rec = (Uh * Sh[None, :]) @ Vh
            rerr = np.linalg.norm(rec - a) / np.linalg.norm(a)
            serr = np.max(np.abs(Sh - ref[:len(Sh)]) / ref[:len(Sh)])
            print(f"{m}x{n} k_sig={k} thr={thr}: sweeps {st[2]:.0f} done {st[3]:.0f} ratio {ref[-1]/ref[0]:.1e} sigma err {serr:.2e} recon {rerr:.2e}", flush=True)
eng.set_option("illcond_thr", 0.005)
