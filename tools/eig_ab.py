"""Development tool: stage and eigen-stage times of vk_compress_batched for the two eigenvector paths."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
cases = [(28, 4, 256, 1024, dict(compressionrank=8)), (28, 4, 256, 1024, dict(decorrelation=0.99)),
         (int(sys.argv[1]) if len(sys.argv) > 1 else 64, 4, 512, 4096, dict(decorrelation=0.99))]
for nbl, nc, m, n, kw in cases:
    A = torch.empty((nbl * nc, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, nc, 0, 2080 if m == 512 else nbl)
    for impl in (1, 0):
        eng.set_option("eigvec_impl", impl)
        eng.set_option("stage_timing", 0)
        for _ in range(2):
            out = eng.compress(A, **kw)
        eng.set_option("stage_timing", 1)
        out = eng.compress(A, **kw)
        torch.cuda.synchronize()
        st, eg = eng.last_stage_ms(), eng.last_eig_ms()
        print(f"{m}x{n} B={nbl*nc} {kw} eigvec_impl={impl}: total {st['total']:.2f} ms  gram {st['gram']:.2f} eig {st['jacobi']:.2f} "
              f"select {st['select']:.2f} factors {st['factors']:.2f} | " + " ".join(f"{k} {v:.2f}" for k, v in eg.items())
              + f" | mean rank {out[3].float().mean().item():.1f} done {int(out[4][:,3].sum().item())}", flush=True)
    del A
