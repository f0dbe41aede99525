"""Dev probe (GPU): energy rule on a parallel-hand-only cube (the reference's default correlation="XX,YY"): leading-pair
path with per-matrix ranks (topk=0) against the full QL path (topk=1) and the Jacobi solver."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
for (nbl, m, n, dec) in [(56, 256, 1024, 0.99), (56, 256, 1024, 0.95), (130, 512, 4096, 0.99)]:
    A = torch.empty((nbl * 2, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, 2)
    eng.set_option("stage_timing", 1)
    for (eig, topk) in ((1, 0), (2, 1), (2, 0)):
        eng.set_option("eig_impl", eig)
        eng.set_option("topk", topk)
        eng.compress(A, decorrelation=dec)
        U, S, Vt, ranks, stats = eng.compress(A, decorrelation=dec)
        torch.cuda.synchronize()
        st = eng.last_stage_ms()
        print(f"{nbl * 2} x {m}x{n} dec={dec} eig_impl={eig} topk={topk}: total {st['total']:.2f} ms eig {st['jacobi']:.2f}  "
              f"mean rank {ranks.float().mean().item():.1f} max {int(ranks.max())}  leading-pair matrices {int((stats[:, 2] == 0).sum())}", flush=True)
    eng.set_option("eig_impl", 0)
    eng.set_option("topk", 0)
    eng.set_option("stage_timing", 0)
