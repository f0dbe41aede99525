"""Dev check of the tcgen05 complex GEMM (V-formation + reconstruction) against the SIMT path / torch."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
dev = "cuda:0"
ok = True
for (B, m, n, k) in [(3, 128, 256, 16), (2, 256, 1024, 64), (2, 512, 1024, 200), (2, 200, 304, 30), (3, 130, 4096, 128), (2, 96, 160, 90)]:
    g = torch.Generator(device=dev).manual_seed(1)
    U = torch.view_as_complex(torch.randn((B, m, k, 2), device=dev, generator=g)).contiguous()
    Vt = torch.view_as_complex(torch.randn((B, k, n, 2), device=dev, generator=g)).contiguous()
    S = torch.rand((B, k), device=dev, generator=g) + 0.5
    ranks = torch.tensor([k, max(1, k // 2), max(1, k - 3)][:B], dtype=torch.int32, device=dev)
    for b in range(B):
        U[b, :, int(ranks[b]):] = 0
        S[b, int(ranks[b]):] = 0
        Vt[b, int(ranks[b]):] = 0
    eng.set_option("gemm_impl", 0)
    out = eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    ref = ((U.to(torch.complex128) * S.to(torch.float64)[:, None, :]) @ Vt.to(torch.complex128))
    err = float((out.to(torch.complex128) - ref).abs().max() / ref.abs().max())
    eng.set_option("gemm_impl", 1)
    out1 = eng.reconstruct(U, S, Vt, ranks)
    err1 = float((out1.to(torch.complex128) - ref).abs().max() / ref.abs().max())
    print(f"recon B={B} m={m} n={n} k={k}: tcgen05 err {err:.2e}  simt err {err1:.2e}", flush=True)
    ok &= err < 3e-6
# full pipeline at large rank: compare compress with gemm_impl 0 vs 1
for (nbl, m, n, kw) in [(2, 256, 1024, dict(decorrelation=0.99)), (1, 512, 4096, dict(decorrelation=0.99)), (2, 128, 512, dict(compressionrank=40))]:
    A = torch.empty((nbl * 4, m, n), dtype=torch.complex64, device=dev)
    eng.synth_fill(A, nbl, 4, nbl_total=8)
    res = {}
    for impl in (0, 1):
        eng.set_option("gemm_impl", impl)
        U, S, Vt, ranks, stats = eng.compress(A, **kw)
        out = eng.reconstruct(U, S, Vt, ranks)
        torch.cuda.synchronize()
        res[impl] = (S.clone(), ranks.clone(), out.clone(), Vt.clone())
    ds = float(((res[0][0] - res[1][0]).abs() / res[1][0].clamp_min(1e-20)).max())
    dr = int((res[0][1] - res[1][1]).abs().max())
    do = float((res[0][2] - res[1][2]).abs().max() / res[1][2].abs().max())
    print(f"compress {m}x{n} {kw}: max rel dS {ds:.2e}  rank diff {dr}  recon diff {do:.2e}  ranks {res[0][1].tolist()[:4]}", flush=True)
    ok &= ds < 1e-5 and dr == 0 and do < 1e-5
print("OK" if ok else "FAIL")
