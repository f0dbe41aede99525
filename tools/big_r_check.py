"""Dev probe (GPU): the Gram path at the upper size limits of the direct eigensolver (r = 1024) and beyond (Jacobi)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
from tests import parity

eng = get_engine(0)
for (B, m, n, kw) in [(2, 1024, 1100, dict(compressionrank=16)), (2, 1024, 1100, dict(decorrelation=0.9)),
                      (2, 700, 2000, dict(compressionrank=40)), (1, 1100, 1200, dict(compressionrank=8)),
                      (2, 1500, 130, dict(decorrelation=0.95))]:
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    U, S, Vt, ranks, stats = eng.compress(A, **kw)
    out = eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    Ah, Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (A, U, S, Vt, ranks, stats))
    for b in range(B):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"{m}x{n} {kw}", **kw)
    print(f"{m}x{n} {kw}: ranks {rk.tolist()} done {st[:, 3].tolist()} iters {st[:, 2].tolist()}  {dt * 1e3:.1f} ms  OK", flush=True)
