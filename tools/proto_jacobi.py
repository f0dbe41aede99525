"""Dev prototype (numpy, complex64): one-sided Jacobi on the Gram matrix G = A A^H.

Checks the design used by the CUDA eigensolver: rotate COLUMNS of Y (Y0 = G) until mutually
orthogonal; then Y = V*Lambda, lambda_i = ||y_i||, u_i = y_i/lambda_i, sigma_i = sqrt(lambda_i).
Not part of the product or the tests.
"""
import sys
import numpy as np


def synth(m, n, nsrc=10, R=15.0, g=1.0, sn=1.0, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(m)[:, None] / m
    nu = np.arange(n)[None, :] / n
    A = np.zeros((m, n), np.complex128)
    for s in range(nsrc):
        rho = rng.uniform(-R, R)
        phi = rng.uniform(0, 2 * np.pi)
        A += np.exp(1j * (2 * np.pi * rho * t * (1 + 0.2 * nu) + phi))
    A *= g
    A += sn * (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
    return A.astype(np.complex64)


def rr_pairs(n):
    """round-robin tournament: n even; yields n-1 rounds of n/2 disjoint pairs"""
    idx = list(range(n))
    for r in range(n - 1):
        yield [(idx[i], idx[n - 1 - i]) for i in range(n // 2)]
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]


def jacobi_onesided(Y, tol, max_sweeps=30, verbose=True):
    Y = Y.astype(np.complex64).copy()
    m = Y.shape[1]
    f32 = np.float32
    for sweep in range(max_sweeps):
        nrot = 0
        maxoff = 0.0
        for pairs in rr_pairs(m):
            ii = np.array([min(p) for p in pairs])
            jj = np.array([max(p) for p in pairs])
            yi = Y[:, ii]
            yj = Y[:, jj]
            a = np.sum((yi.real ** 2 + yi.imag ** 2), axis=0, dtype=f32)
            b = np.sum((yj.real ** 2 + yj.imag ** 2), axis=0, dtype=f32)
            z = np.sum(np.conj(yi) * yj, axis=0).astype(np.complex64)
            az = np.abs(z).astype(f32)
            denom = np.sqrt(a * b)
            rel = np.where(denom > 0, az / np.maximum(denom, f32(1e-38)), 0)
            maxoff = max(maxoff, float(rel.max()))
            rot = (rel > tol) & (az > 0)
            nrot += int(rot.sum())
            # 2x2 hermitian [[a, z],[conj z, b]] -> rotation
            azs = np.where(rot, az, f32(1))
            ph = np.where(rot, z / azs, 1).astype(np.complex64)
            tau = ((b - a) / (2 * azs)).astype(f32)
            t = (np.sign(tau) / (np.abs(tau) + np.sqrt(1 + tau * tau))).astype(f32)
            t = np.where(tau == 0, f32(1), t)
            c = (1 / np.sqrt(1 + t * t)).astype(f32)
            s = (c * t).astype(f32)
            c = np.where(rot, c, f32(1))
            s = np.where(rot, s, f32(0))
            w = (s * ph).astype(np.complex64)      # complex sine
            ni = c * yi - np.conj(w) * yj
            nj = w * yi + c * yj
            Y[:, ii] = ni.astype(np.complex64)
            Y[:, jj] = nj.astype(np.complex64)
        if verbose:
            print(f"  sweep {sweep}: rotations {nrot}, max rel off {maxoff:.3e}")
        if nrot == 0:
            break
    return Y, sweep + 1


if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    g = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    A = synth(m, n, g=g)
    U0, S0, Vt0 = np.linalg.svd(A, full_matrices=False)
    S64 = np.linalg.svd(A.astype(np.complex128), compute_uv=False)
    G = (A.astype(np.complex128) @ A.astype(np.complex128).conj().T).astype(np.complex64)
    tol = np.float32(np.sqrt(m) * 6e-8)
    Y, ns = jacobi_onesided(G, tol)
    lam = np.sqrt(np.sum(np.abs(Y.astype(np.complex128)) ** 2, axis=0))
    order = np.argsort(-lam)
    lam = lam[order]
    U = (Y[:, order] / lam[None, :]).astype(np.complex64)
    sig = np.sqrt(lam)
    print("sweeps", ns)
    rel = np.abs(sig - S64) / S64
    print("sigma rel err (sqrt lambda): max over all", rel.max(), " over sigma>0.02*s1:", rel[S64 > 0.02 * S64[0]].max())
    B = U.conj().T.astype(np.complex64) @ A
    sref = np.sqrt(np.sum(np.abs(B.astype(np.complex128)) ** 2, axis=1))
    rel2 = np.abs(sref - S64) / S64
    print("sigma rel err (refined ||A^H u||): max", rel2.max())
    print("oracle f32 svd rel err vs f64:", (np.abs(S0 - S64) / S64).max())
    orth = np.abs(U.conj().T @ U - np.eye(m)).max()
    print("U orth err", orth)
    for k in (1, 8, m // 4, m // 2):
        rec = (U[:, :k] @ B[:k]).astype(np.complex64)
        e = np.linalg.norm(A - rec)
        rec0 = ((U0[:, :k] * S0[None, :k]) @ Vt0[:k]).astype(np.complex64)
        e0 = np.linalg.norm(A - rec0)
        print(f"k={k}: err ours {e:.6f} ref {e0:.6f} rel diff {(e - e0) / e0:.2e}")
