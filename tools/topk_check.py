"""Dev: fixed-rank fast path (subspace iteration) vs full Jacobi — time, how many matrices it solved, agreement."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
for (nbl, m, n) in [(28, 256, 1024), (16, 512, 4096)]:
    A = torch.empty((nbl * 4, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, 4, nbl_total=28)
    for k in (1, 2, 4, 8):
        res = {}
        for topk in (1, 0):
            eng.set_option("topk", topk)
            eng.set_option("stage_timing", 1)
            for _ in range(2):
                U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
                torch.cuda.synchronize()
            ms = eng.last_stage_ms()["jacobi"]
            out = eng.reconstruct(U, S, Vt, ranks)
            torch.cuda.synchronize()
            res[topk] = (S.clone(), out.clone(), ms, stats[:, 2].clone())
        eng.set_option("stage_timing", 0)
        dS = float(((res[0][0] - res[1][0]).abs() / res[1][0]).max())
        dO = float((res[0][1] - res[1][1]).abs().max() / res[1][1].abs().max())
        e1 = float((A - res[1][1]).abs().pow(2).sum().sqrt())
        e0 = float((A - res[0][1]).abs().pow(2).sum().sqrt())
        par = res[0][3].view(-1, 4)[:, [0, 3]].float().mean().item()
        crs = res[0][3].view(-1, 4)[:, [1, 2]].float().mean().item()
        print(f"{m}x{n} k={k}: eig stage full {res[1][2]:.2f} ms, fast {res[0][2]:.2f} ms | max rel dS {dS:.1e} recon diff {dO:.1e} "
              f"err ratio {e0 / e1 - 1:+.1e} | mean its/sweeps parallel-hand {par:.1f} cross-hand {crs:.1f}", flush=True)
