"""Prototype (numpy, float32 arithmetic): all eigenvectors of the real symmetric tridiagonal T by bisection + twisted
factorisation (no reorthogonalisation), then one or two Newton-Schulz steps Z <- Z (I - E/2), E = Z^T Z - I, done as
GEMMs. Measures what the CUDA path will see: max |E| before / after, residuals, and the effect on the subspace.
Development tool only; not part of the product path."""
import sys
import numpy as np
import scipy.linalg as sl

sys.path.insert(0, ".")
from oracle.synth_np import synth_cube  # noqa: E402

f32 = np.float32


def tridiag64(G):
    H, Q = sl.hessenberg(G.astype(np.complex128), calc_q=True)
    d = H.diagonal().real.copy()
    sub = np.array([H[i + 1, i] for i in range(len(d) - 1)])
    return d.astype(f32), np.abs(sub).astype(f32)


def sturm_counts(d, e2, x, pivmin):
    """number of eigenvalues < x[t] for every t (vectorised over t), float32"""
    q = d[0] - x
    q = np.where(np.abs(q) < pivmin, -pivmin, q).astype(f32)
    c = (q < 0).astype(np.int32)
    for i in range(1, len(d)):
        q = (d[i] - x - e2[i - 1] / q).astype(f32)
        q = np.where(np.abs(q) < pivmin, -pivmin, q).astype(f32)
        c += q < 0
    return c


def bisect_all(d, e):
    r = len(d)
    e2 = (e * e).astype(f32)
    rad = np.zeros(r, f32)
    rad[:-1] += np.abs(e)
    rad[1:] += np.abs(e)
    lo, hi = f32((d - rad).min()), f32((d + rad).max())
    scale = max(abs(lo), abs(hi))
    pivmin = f32(max(1e-30, 1e-14 * scale * scale))
    idx = np.arange(r)            # ascending index
    a = np.full(r, lo - 1e-6 * scale, f32)
    c = np.full(r, hi + 1e-6 * scale, f32)
    for _ in range(48):
        mid = (f32(0.5) * (a + c)).astype(f32)
        cnt = sturm_counts(d, e2, mid, pivmin)
        up = cnt > idx
        c = np.where(up, mid, c)
        a = np.where(up, a, mid)
    return (f32(0.5) * (a + c)).astype(f32)     # ascending


def twisted_all(d, e, lam):
    """one vector per eigenvalue (columns of Z), float32, vectorised over eigenvalues"""
    r = len(d)
    pivmin = np.maximum(1e-30, 1e-14 * lam * lam).astype(f32)
    dm = np.zeros((r, r), f32)    # [i][t]
    q = (d[r - 1] - lam).astype(f32)
    q = np.where(np.abs(q) < pivmin, -pivmin, q).astype(f32)
    dm[r - 1] = q
    for i in range(r - 2, -1, -1):
        q = (d[i] - lam - e[i] * e[i] / q).astype(f32)
        q = np.where(np.abs(q) < pivmin, -pivmin, q).astype(f32)
        dm[i] = q
    dp = np.zeros((r, r), f32)
    p = (d[0] - lam).astype(f32)
    p = np.where(np.abs(p) < pivmin, -pivmin, p).astype(f32)
    dp[0] = p
    best = np.abs(dm[0]).copy()
    kt = np.zeros(r, np.int64)
    for i in range(1, r):
        p = (d[i] - lam - e[i - 1] * e[i - 1] / p).astype(f32)
        p = np.where(np.abs(p) < pivmin, -pivmin, p).astype(f32)
        dp[i] = p
        gam = np.abs(p + dm[i] - (d[i] - lam)).astype(f32)
        upd = gam < best
        best = np.where(upd, gam, best)
        kt = np.where(upd, i, kt)
    Z = np.zeros((r, r), f32)
    for t in range(r):
        k = kt[t]
        z = np.zeros(r, f32)
        z[k] = 1
        zi = f32(1)
        for i in range(k - 1, -1, -1):
            zi = f32(-(e[i] / dp[i, t]) * zi)
            z[i] = zi
        zi = f32(1)
        for i in range(k + 1, r):
            zi = f32(-(e[i - 1] / dm[i, t]) * zi)
            z[i] = zi
        with np.errstate(over="ignore", invalid="ignore"):
            nrm = np.sqrt(np.sum(z.astype(np.float64) ** 2))
        Z[:, t] = (z / nrm).astype(f32) if np.isfinite(nrm) and nrm > 0 else np.nan
    return Z


def report(name, d, e):
    r = len(d)
    T = np.diag(d.astype(np.float64)) + np.diag(e.astype(np.float64), 1) + np.diag(e.astype(np.float64), -1)
    lam_ref, Zref = np.linalg.eigh(T)
    lam = bisect_all(d, e)
    Z = twisted_all(d, e, lam)
    if not np.isfinite(Z).all():
        print(f"{name}: non-finite vectors ({np.isnan(Z).any(axis=0).sum()} columns)")
        return
    sc = np.abs(lam_ref).max()
    res = np.abs(T @ Z - Z * lam[None, :]).max() / sc
    E = (Z.T.astype(f32) @ Z.astype(f32) - np.eye(r, dtype=f32)).astype(f32)
    e0 = np.abs(E).max()
    n2 = np.linalg.norm(E.astype(np.float64), 2)
    Z1 = (Z - f32(0.5) * (Z @ E)).astype(f32)
    E1 = (Z1.T @ Z1 - np.eye(r, dtype=f32)).astype(f32)
    Z2 = (Z1 - f32(0.5) * (Z1 @ E1)).astype(f32)
    E2 = (Z2.T @ Z2 - np.eye(r, dtype=f32)).astype(f32)
    res2 = np.abs(T @ Z2 - Z2 * lam[None, :]).max() / sc
    gaps = np.diff(lam_ref) / sc
    print(f"{name}: r={r} lam err {np.abs(lam - lam_ref).max() / sc:.1e}  min gap {gaps.min():.1e}  resid {res:.1e}  "
          f"max|E| {e0:.1e} ||E||2 {n2:.1e}  after NS1 {np.abs(E1).max():.1e}  NS2 {np.abs(E2).max():.1e}  resid(NS2) {res2:.1e}")


def main():
    rng = np.random.default_rng(1)
    for (m, n) in [(256, 1024), (512, 2048)]:
        cube = synth_cube(2, 4, m, n, nbl_total=8)
        for b, label in [(0, "parallel"), (1, "cross"), (4, "parallel-long")]:
            A = cube[b].astype(np.complex128)
            G = A @ A.conj().T
            G /= np.trace(G).real / m
            d, e = tridiag64(G)
            report(f"{label} {m}x{n}", d, e)
    # exactly low rank (rank 20 of 256): most eigenvalues are round-off zeros
    m, n = 256, 1024
    L = (rng.standard_normal((m, 20)) + 1j * rng.standard_normal((m, 20)))
    Rr = (rng.standard_normal((20, n)) + 1j * rng.standard_normal((20, n)))
    A = (L @ Rr).astype(np.complex64).astype(np.complex128)
    G = (A @ A.conj().T).astype(np.complex64).astype(np.complex128)
    G /= np.trace(G).real / m
    d, e = tridiag64(G)
    report("low-rank 20 of 256", d, e)
    # multiple eigenvalues: G = Q diag(1,1,1,1,0.5,...) Q^H
    lam = np.concatenate([np.ones(8), np.full(8, 0.5), np.linspace(0.4, 0.01, m - 16)])
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    G = (Q * lam[None, :]) @ Q.conj().T
    G = G.astype(np.complex64).astype(np.complex128)
    d, e = tridiag64(G)
    report("multiple eigenvalues", d, e)


if __name__ == "__main__":
    main()
