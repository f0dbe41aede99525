"""Dev probe (GPU): reconstruction time of the 1040-matrix MeerKAT shard from its own factors (ragged ranks)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
A = torch.empty((1040, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, 260, 4)
U, S, Vt, ranks, stats = eng.compress(A, decorrelation=0.99)
out = torch.empty_like(A)
for rep in range(3):
    eng.reconstruct(U, S, Vt, ranks, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.reconstruct(U, S, Vt, ranks, out=out)
    e1.record(); e1.synchronize()
    print("reconstruct", e0.elapsed_time(e1) / 3, "ms; mean rank", float(ranks.float().mean()), flush=True)
