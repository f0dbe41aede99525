"""Dev probe (GPU): the direct eigensolver (eig_impl=2: tridiagonalisation + implicit QL) against numpy and against the
Jacobi solver, through vk_eigh_jacobi_batched and through the whole compress path. Usage: eig_check.py [B m n]..."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from visco_b200.engine import get_engine  # noqa: E402


def stage_check(eng, B, m, n):
    dev = torch.device("cuda:0")
    A = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
    eng.synth_fill(A, B, 1)
    G = eng.gram(A)
    r = G.shape[1]
    G64 = G[:2].cpu().numpy().astype(np.complex128)
    for impl in (1, 2):
        eng.set_option("eig_impl", impl)
        W = G.clone()
        torch.cuda.synchronize()
        lam, info = eng.eigh_jacobi(W)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            W = G.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lam, info = eng.eigh_jacobi(W)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        lam_h = lam.cpu().numpy()
        info_h = info.cpu().numpy()
        Wh = W[:2].cpu().numpy().astype(np.complex128)
        worst = dict(lam=0.0, orth=0.0, res=0.0)
        for b in range(2):
            ref = np.linalg.eigvalsh(G64[b])[::-1]
            worst["lam"] = max(worst["lam"], np.abs(lam_h[b] - ref).max() / ref.max())
            nrm = np.linalg.norm(Wh[b], axis=1)
            V = (Wh[b] / np.maximum(nrm, 1e-300)[:, None]).T  # columns = eigenvectors of G
            worst["orth"] = max(worst["orth"], np.abs(V.conj().T @ V - np.eye(r)).max())
            # G V = V diag(lambda): use the norms (in trace-normalised units) rescaled to G
            scale = np.trace(G64[b]).real / r
            worst["res"] = max(worst["res"], np.abs(G64[b] @ V - V * (nrm * scale)[None, :]).max() / ref.max())
        print(f"  eig_impl={impl}: {min(ts):8.2f} ms  sweeps/iters {info_h[:, 0].mean():.1f} done {int(info_h[:, 1].min())}"
              f"  lam err {worst['lam']:.2e}  orth {worst['orth']:.2e}  resid {worst['res']:.2e}")
    eng.set_option("eig_impl", 1)


def path_check(eng, B, m, n, **kw):
    dev = torch.device("cuda:0")
    A = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
    eng.synth_fill(A, B // 4 if B % 4 == 0 else B, 4 if B % 4 == 0 else 1)
    eng.set_option("stage_timing", 1)
    outs = {}
    for impl in (1, 2):
        eng.set_option("eig_impl", impl)
        res = eng.compress(A, **kw)
        torch.cuda.synchronize()
        res = eng.compress(A, **kw)
        torch.cuda.synchronize()
        st = eng.last_stage_ms()
        U, S, Vt, ranks, stats = res
        out = eng.reconstruct(U, S, Vt, ranks)
        err = (torch.linalg.norm((out - A).reshape(B, -1), dim=1) / torch.linalg.norm(A.reshape(B, -1), dim=1))
        outs[impl] = (S.cpu().numpy(), ranks.cpu().numpy(), err.cpu().numpy())
        print(f"  path eig_impl={impl} {kw}: total {st['total']:.2f} ms (gram {st['gram']:.2f} eig {st['jacobi']:.2f} "
              f"select {st['select']:.2f} factors {st['factors']:.2f})  err max {err.max().item():.3e}")
    S1, r1, e1 = outs[1]
    S2, r2, e2 = outs[2]
    k = min(S1.shape[1], 16)
    print(f"  ranks equal: {np.array_equal(r1, r2)}  S rel diff (top {k}) {np.abs(S1[:, :k] - S2[:, :k]).max() / S1.max():.2e}"
          f"  err diff {np.abs(e1 - e2).max():.2e}")
    eng.set_option("eig_impl", 1)
    eng.set_option("stage_timing", 0)


def main():
    eng = get_engine(0)
    shapes = [(8, 100, 300), (112, 256, 1024), (64, 512, 4096)]
    if len(sys.argv) > 3:
        a = [int(x) for x in sys.argv[1:]]
        shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
    for (B, m, n) in shapes:
        print(f"B={B} {m}x{n}")
        stage_check(eng, B, m, n)
        path_check(eng, B, m, n, compressionrank=8)
        path_check(eng, B, m, n, decorrelation=0.99)


main()
