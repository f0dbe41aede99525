"""Development tool: one compress + reconstruct of B MeerKAT-shaped matrices (energy rule) for ncu launch lists."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
A = torch.empty((B, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B // 4, 4, nbl_total=2080, bl_offset=500)
for _ in range(2):
    U, S, Vt, ranks, stats = eng.compress(A, decorrelation=0.99)
    out = eng.reconstruct(U, S, Vt, ranks)
torch.cuda.synchronize()
print("ok", float(ranks.float().mean()))
