"""One compress + reconstruct of the KAT-7 cube (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
nbl, ncorr, m, n, kw = {"c2": (28, 4, 256, 1024, dict(compressionrank=8)), "c3": (4, 4, 512, 4096, dict(decorrelation=0.99)),
                        "c4": (2080, 4, 64, 64, dict(compressionrank=8))}[wl]
A = torch.empty((nbl * ncorr, m, n), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, nbl, ncorr, nbl_total=max(nbl, 28))
U, S, Vt, ranks, stats = eng.compress(A, **kw)
out = eng.reconstruct(U, S, Vt, ranks)
torch.cuda.synchronize()
print("ok", float(stats[:, 2].mean()))
