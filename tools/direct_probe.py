"""Dev probe (GPU): the Gram-free redo path forced on every matrix (illcond_thr = 1) at several sizes, against the oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests import parity
from visco_b200.engine import get_engine

eng = get_engine(0)
rng = np.random.default_rng(3)
for (m, n) in [(100, 150), (200, 260), (300, 320), (442, 448), (520, 540), (586, 627), (130, 700), (700, 130)]:
    A = (rng.standard_normal((2, m, n)) + 1j * rng.standard_normal((2, m, n))).astype(np.complex64)
    Ad = torch.from_numpy(A).cuda()
    for thr in (0.0, 1.0):
        eng.set_option("illcond_thr", thr)
        try:
            U, S, Vt, ranks, stats = eng.compress(Ad)
            torch.cuda.synchronize()
            Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (U, S, Vt, ranks, stats))
            k = int(rk[0])
            ref = np.linalg.svd(A[0].astype(np.complex128), compute_uv=False)
            serr = np.max(np.abs(Sh[0, :k] - ref[:k]) / ref[:k])
            orth_u = np.abs(Uh[0].conj().T @ Uh[0] - np.eye(k)).max()
            orth_v = np.abs(Vh[0] @ Vh[0].conj().T - np.eye(k)).max()
            rec = (Uh[0] * Sh[0][None, :]) @ Vh[0]
            rerr = np.linalg.norm(rec - A[0]) / np.linalg.norm(A[0])
            print(f"{m}x{n} thr={thr}: sweeps {st[:, 2]} done {st[:, 3]} sigma err {serr:.2e} (min ratio {ref[k-1]/ref[0]:.1e}) "
                  f"orthU {orth_u:.1e} orthV {orth_v:.1e} recon {rerr:.1e}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"{m}x{n} thr={thr}: {type(ex).__name__}: {str(ex)[:200]}", flush=True)
eng.set_option("illcond_thr", 0.005)
