"""Prototype (numpy, float32 arithmetic): Hermitian tridiagonalisation + implicit QL with recorded plane rotations,
to measure rotation counts, level-schedule depth and float32 accuracy before committing to CUDA kernels.
Development tool only; not part of the product path."""
import sys
import time
import numpy as np

sys.path.insert(0, ".")
from oracle.synth_np import synth_cube  # noqa: E402

f32 = np.float32
c64 = np.complex64


def tridiag(G):
    """G Hermitian complex64 (r x r). Returns d, e (real float32), Q (complex64) with Q^H G Q = T (real sym tridiag)."""
    r = G.shape[0]
    A = G.copy()
    Q = np.eye(r, dtype=c64)
    vs = []
    for j in range(r - 2):
        x = A[j + 1:, j].copy()
        alpha = x[0]
        xn = f32(np.sqrt(np.sum((x.real.astype(f32)) ** 2 + (x.imag.astype(f32)) ** 2, dtype=f32)))
        if xn == 0:
            vs.append(None)
            continue
        ph = alpha / abs(alpha) if abs(alpha) > 0 else c64(1)
        v = x.copy()
        v[0] = alpha + ph * xn
        vn2 = f32(np.sum(v.real ** 2 + v.imag ** 2, dtype=f32))
        tau = f32(2) / vn2
        # two-sided update of trailing block: A <- H A H, H = I - tau v v^H
        At = A[j + 1:, j + 1:]
        p = (tau * (At @ v)).astype(c64)
        K = (tau / f32(2)) * np.vdot(v, p)
        w = (p - K * v).astype(c64)
        At -= np.outer(v, w.conj()) + np.outer(w, v.conj())
        A[j + 1:, j] = 0
        A[j + 1, j] = -ph * xn
        A[j, j + 1:] = A[j + 1:, j].conj()
        vs.append((v, tau))
    d = A.diagonal().real.astype(f32).copy()
    esub = np.array([A[i + 1, i] for i in range(r - 1)], dtype=c64)
    # Q = H_0 H_1 ... ; form by backward accumulation
    for j in range(r - 3, -1, -1):
        if vs[j] is None:
            continue
        v, tau = vs[j]
        Qt = Q[j + 1:, j + 1:]
        Qt -= np.outer(tau * v, (v.conj() @ Qt)).astype(c64)
    # phases: make sub-diagonal real non-negative: T_real = D^H T D
    ph = np.ones(r, dtype=c64)
    for i in range(r - 1):
        a = abs(esub[i])
        ph[i + 1] = ph[i] * (esub[i] / a if a > 0 else 1)
    e = np.abs(esub).astype(f32)
    Q = (Q * ph[None, :]).astype(c64)
    return d, e, Q


def tql_rotations(d, e, maxit=60):
    """Implicit QL (tqli), float32; records rotations (col i, c, s) acting on columns (i, i+1) of the vector matrix."""
    n = len(d)
    d = d.astype(f32).copy()
    e = np.concatenate([e.astype(f32), [f32(0)]])
    rots = []
    sweeps = 0
    eps = np.finfo(f32).eps
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) <= eps * dd:
                    break
                m += 1
            if m == l:
                break
            it += 1
            assert it <= maxit
            sweeps += 1
            g = (d[l + 1] - d[l]) / (f32(2) * e[l])
            rr = f32(np.hypot(g, f32(1)))
            g = d[m] - d[l] + e[l] / (g + (rr if g >= 0 else -rr))
            s = f32(1)
            c = f32(1)
            p = f32(0)
            i = m - 1
            early = False
            while i >= l:
                f = s * e[i]
                b = c * e[i]
                rr = f32(np.hypot(f, g))
                e[i + 1] = rr
                if rr == 0:
                    d[i + 1] -= p
                    e[m] = 0
                    early = True
                    break
                s = f / rr
                c = g / rr
                g = d[i + 1] - p
                rr = (d[i] - g) * s + f32(2) * c * b
                p = s * rr
                d[i + 1] = g + p
                g = c * rr - b
                rots.append((i, c, s))
                i -= 1
            if early:
                continue
            d[l] -= p
            e[l] = g
            e[m] = 0
    return d, rots, sweeps


def apply_rots(X, rots):
    X = X.copy()
    for (i, c, s) in rots:
        a = X[:, i].copy()
        b = X[:, i + 1].copy()
        X[:, i + 1] = s * a + c * b
        X[:, i] = c * a - s * b
    return X


def levels(rots, n):
    last = np.zeros(n + 1, dtype=np.int64)
    mx = 0
    for (i, c, s) in rots:
        lv = max(last[i], last[i + 1]) + 1
        last[i] = last[i + 1] = lv
        mx = max(mx, lv)
    return mx


def main():
    m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 1024)
    A = synth_cube(1, 1, m, n)[0]
    G = (A @ A.conj().T).astype(c64)
    G = (G / f32(np.trace(G).real / m)).astype(c64)
    t0 = time.time()
    d, e, Q = tridiag(G)
    t1 = time.time()
    T = np.diag(d.astype(np.float64)) + np.diag(e.astype(np.float64), 1) + np.diag(e.astype(np.float64), -1)
    Q64 = Q.astype(np.complex128)
    print("tridiag %.1fs  |Q^H G Q - T|/|G| = %.2e   |Q^H Q - I| = %.2e" % (
        t1 - t0, np.abs(Q64.conj().T @ G.astype(np.complex128) @ Q64 - T).max() / np.abs(G).max(),
        np.abs(Q64.conj().T @ Q64 - np.eye(m)).max()))
    lam, rots, sweeps = tql_rotations(d, e)
    t2 = time.time()
    print("tql %.1fs rotations %d (%.2f r^2) sweeps %d  levels %d" % (t2 - t1, len(rots), len(rots) / m / m, sweeps,
                                                                      levels(rots, m)))
    V = apply_rots(Q, rots)
    V64 = V.astype(np.complex128)
    G64 = G.astype(np.complex128)
    ref = np.linalg.eigvalsh(G64)
    print("eigenvalue err (abs/max) %.2e" % (np.abs(np.sort(lam.astype(np.float64)) - ref).max() / ref.max()))
    print("|V^H V - I| max %.2e   |G V - V L| max/|G| %.2e" % (
        np.abs(V64.conj().T @ V64 - np.eye(m)).max(),
        np.abs(G64 @ V64 - V64 * lam.astype(np.float64)[None, :]).max() / ref.max()))
    # off-diagonal of V^H G V relative to sqrt(l_i l_j): what a Jacobi polish would see
    H = V64.conj().T @ G64 @ V64
    dd = np.sqrt(np.abs(np.outer(H.diagonal().real, H.diagonal().real)))
    off = np.abs(H - np.diag(H.diagonal())) / dd
    print("scaled off-diagonal max %.2e  (Jacobi on G-columns sees the squared-spectrum analogue)" % off.max())
    # full-rank reconstruction through the pipeline: U = V, B = U^H A, out = U B
    U = V
    B = (U.conj().T @ A).astype(c64)
    out = (U @ B).astype(c64)
    print("full-rank reconstruction rel err %.2e" % (np.linalg.norm(out - A) / np.linalg.norm(A)))


main()
