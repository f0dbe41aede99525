"""Dev probe (GPU): fixed-rank path on matrices with exactly / nearly multiple leading singular values."""
import sys
import numpy as np
sys.path.insert(0, ".")
from visco_b200.compress_ms import apply_svd
from visco_b200.decompress_ms import reconstruct_vis

rng = np.random.default_rng(3)
m, n = 128, 256
Q1, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
Q2, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
cases = {
    "scaled unitary rows (all sigma equal)": np.ones(m),
    "three equal leading values": np.concatenate([[5, 5, 5, 2, 1], 0.01 * np.ones(m - 5)]),
    "pair 1e-6 apart": np.concatenate([[5, 5 * (1 - 1e-6), 3, 2, 1], 0.01 * rng.random(m - 5)]),
    "pair 1e-4 apart": np.concatenate([[5, 5 * (1 - 1e-4), 3, 2, 1], 0.01 * rng.random(m - 5)]),
    "rank 2, k = 4": np.concatenate([[3, 1], np.zeros(m - 2)]),
}
for name, sv in cases.items():
    A = ((Q1 * sv[None, :]) @ Q2[:m]).astype(np.complex64)
    k = 4
    U, S, Vt = apply_svd(A, compressionrank=k)
    rec = reconstruct_vis(U, S, Vt)
    s_ref = np.sort(sv)[::-1][:k]
    best = np.sqrt(np.sum(np.sort(sv)[::-1][k:] ** 2))
    err = np.linalg.norm(A - rec)
    orth = np.abs(U.conj().T @ U - np.eye(k)).max()
    print(f"{name:40s} S {np.array2string(S, precision=5)} ref {np.array2string(s_ref, precision=5)}  err {err:.4e} (optimal {best:.4e})  |U^H U - I| {orth:.1e}")
