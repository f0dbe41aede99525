"""GPU parity tests: every call goes through the C ABI (libvisco_b200.so); the oracle is only the checker."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from visco_b200.engine import get_engine
    return get_engine(0)


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def test_library_is_the_cuda_one(eng):
    assert eng.lib.vk_version().decode().endswith("(sm_100a)")
    assert eng.launch_count >= 0


# ---------------------------------------------------------------------------------------- golden fixtures
def test_golden_full_rank_singular_values(golden, golden_cases):
    from visco_b200.compress_ms import apply_svd
    for c in golden_cases:
        a = golden[f"{c}/A"]
        U, S, Vt = apply_svd(a)
        parity.check_factors(a, U, S, Vt, len(S), label=c)
        s_ref = golden[f"{c}/S"]
        big = s_ref > 1e-3 * s_ref[0]
        np.testing.assert_allclose(S[big], s_ref[big], rtol=1e-4, err_msg=c)


def test_golden_energy_rule(golden, golden_cases):
    from visco_b200.compress_ms import apply_svd, find_n_decorrelation
    decs = golden["decs"]
    for c in golden_cases:
        a = golden[f"{c}/A"]
        s_ref = golden[f"{c}/S"]
        for d, n_ref in zip(decs, golden[f"{c}/n_dec"]):
            # the device restatement of find_n_decorrelation on the REFERENCE's singular values: exact match
            assert find_n_decorrelation(s_ref, float(d)) == int(n_ref), (c, d)
        for d in (0.9, 0.99):
            U, S, Vt = apply_svd(a, decorrelation=float(d))
            parity.check_factors(a, U, S, Vt, len(S), decorrelation=float(d), label=f"{c}@{d}")


def test_golden_fixed_rank_and_reconstruction(golden, golden_cases):
    from visco_b200.compress_ms import apply_svd
    from visco_b200.decompress_ms import reconstruct_vis
    for c in golden_cases:
        a = golden[f"{c}/A"]
        for k, e_ref in zip(golden[f"{c}/ks"], golden[f"{c}/recon_err"]):
            U, S, Vt = apply_svd(a, decorrelation=0.9, compressionrank=int(k))     # fixed rank wins
            assert len(S) == k
            parity.check_factors(a, U, S, Vt, int(k), compressionrank=int(k), label=f"{c}/k{k}")
            rec = reconstruct_vis(U, S, Vt)
            parity.check_reconstruction(U, S, Vt, rec, label=f"{c}/k{k}")
            e = np.linalg.norm(a.astype(np.complex128) - rec.astype(np.complex128))
            assert abs(e - e_ref) <= 1e-5 * e_ref + 3e-6 * np.linalg.norm(a), (c, k, e, e_ref)
            rec2 = reconstruct_vis(U, S.reshape(-1, 1), Vt)                          # (k, 1) form accepted
            np.testing.assert_array_equal(rec, rec2)


def test_cross_reading_reference_factors(golden):
    """decompress side reads factors the REFERENCE produced (oracle = reference restatement)."""
    from oracle import visco_oracle as vo
    from visco_b200.decompress_ms import reconstruct_vis, unstack_vis
    a = golden["ms_bl12_diag/A"]
    u, s, vt = vo.ref_apply_svd(a, compressionrank=4)
    rec = reconstruct_vis(u, s, vt)
    parity.check_reconstruction(u, s, vt, rec)
    parts = unstack_vis(rec, 360)
    assert len(parts) == 2 and parts[0].shape == (360, 16)
    np.testing.assert_allclose(parts[0], golden["ms_bl12_diag/unstack0"], atol=3e-5 * np.abs(a).max())
    np.testing.assert_allclose(parts[1], golden["ms_bl12_diag/unstack1"], atol=3e-5 * np.abs(a).max())


# ---------------------------------------------------------------------------------------- synthetic cubes
def _device_cube(eng, torch, nbl, ncorr, m, n, nbl_total=None, bl_offset=0):
    A = torch.empty((nbl * ncorr, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, ncorr, bl_offset=bl_offset, nbl_total=nbl_total or nbl)
    return A


@pytest.mark.parametrize("shape,kw", [
    ((256, 1024), dict(compressionrank=8)),          # KAT-7 config (C2) shape, Gram path
    ((128, 512), dict(decorrelation=0.99)),          # energy rule, Gram path, near-full rank on cross hands
    ((64, 64), dict(decorrelation=0.95)),            # small-matrix regime (C4), direct path
    ((64, 64), dict(compressionrank=5)),
    ((360, 16), dict(decorrelation=0.9)),            # sample-MS shape (m > n), direct path
    ((96, 40), dict(compressionrank=3)),
    ((200, 72), dict(decorrelation=0.97)),           # m > n on the Gram path (r > 64)
    ((70, 300), dict()),                             # full rank, odd sizes
])
def test_synthetic_cube_against_oracle(eng, torch, shape, kw):
    m, n = shape
    nbl, ncorr = 3, 4
    A = _device_cube(eng, torch, nbl, ncorr, m, n, nbl_total=8, bl_offset=2)
    U, S, Vt, ranks, stats = eng.compress(A, **kw)
    out = eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    Ah, Uh, Sh, Vh, rk, st, oh = (x.cpu().numpy() for x in (A, U, S, Vt, ranks, stats, out))
    assert np.all(st[:, 3] == 1), "Jacobi did not converge"
    for b in range(nbl * ncorr):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"{shape}{kw} b={b}", **kw)
        assert not Uh[b, :, k:].any() and not Sh[b, k:].any() and not Vh[b, k:].any(), "padding must be zero"
        parity.check_reconstruction(Uh[b, :, :k], Sh[b, :k], Vh[b, :k], oh[b], label=f"recon b={b}")
        # stats: total and retained energy
        tot = float(np.sum(np.abs(Ah[b].astype(np.complex128)) ** 2))
        assert abs(st[b, 0] - tot) <= 1e-4 * tot
        assert abs(st[b, 1] - float(np.sum(Sh[b, :k].astype(np.float64) ** 2))) <= 1e-4 * tot


def test_host_api_matches_device_api(eng, torch):
    A = _device_cube(eng, torch, 2, 4, 64, 96)
    U, S, Vt, ranks, _ = eng.compress(A, compressionrank=6)
    torch.cuda.synchronize()
    Uh, Sh, Vh, rh, _ = eng.compress_host(A.cpu().numpy(), compressionrank=6)
    np.testing.assert_array_equal(rh, ranks.cpu().numpy())
    np.testing.assert_allclose(Sh, S.cpu().numpy(), rtol=1e-6)
    out_h = eng.reconstruct_host(Uh, Sh, Vh, rh)
    out_d = eng.reconstruct(U, S, Vt, ranks).cpu().numpy()
    np.testing.assert_allclose(out_h, out_d, atol=1e-5 * np.abs(out_d).max())


# ---------------------------------------------------------------------------------------- stages
def test_gram_stage_simt(eng, torch):
    for (m, n) in [(96, 200), (200, 72), (130, 130)]:
        A = _device_cube(eng, torch, 2, 2, m, n)
        W = eng.gram(A, impl=1).cpu().numpy()
        a = A.cpu().numpy().astype(np.complex128)
        for b in range(A.shape[0]):
            G = a[b] @ a[b].conj().T if m <= n else a[b].conj().T @ a[b]
            np.testing.assert_allclose(W[b], G.T, atol=2e-6 * np.abs(G).max())


def test_gram_stage_tcgen05(eng, torch):
    """tcgen05 + TMA Gram kernel (3xTF32, two-level accumulation) against an fp64 Gram; includes ragged tiles
    (m not a multiple of 128) and a K tail (2n not a multiple of 32)."""
    for (B, m, n) in [(2, 128, 256), (3, 256, 1024), (2, 200, 300), (2, 512, 1024), (2, 130, 130), (1, 96, 2050)]:
        assert eng.gram_uses_tcgen05(m, n)
        A = _device_cube(eng, torch, B, 1, m, n, nbl_total=8)
        W = eng.gram(A, impl=2).cpu().numpy()
        a = A.cpu().numpy().astype(np.complex128)
        G = np.einsum("btv,biv->bit", a, a.conj())
        assert np.abs(W - G).max() <= 3e-6 * np.abs(G).max(), (m, n, np.abs(W - G).max() / np.abs(G).max())
        # Hermitian: off-diagonal tiles are mirrored exactly, diagonal tiles agree to accumulation-order rounding
        np.testing.assert_allclose(W, W.conj().transpose(0, 2, 1), atol=1e-6 * np.abs(G).max())
    assert not eng.gram_uses_tcgen05(64, 4096) and not eng.gram_uses_tcgen05(300, 200)


@pytest.mark.parametrize("impl,m,n", [(1, 160, 400), (2, 160, 400), (2, 65, 97), (2, 100, 64), (2, 333, 700),
                                      (2, 2, 5), (2, 3, 9), (2, 600, 1200)])
def test_eigh_stage(eng, torch, impl, m, n):
    """vk_eigh_jacobi_batched with the cyclic Jacobi solver (eig_impl=1) and the direct one (eig_impl=2:
    tridiagonalisation + implicit QL): eigenvalues, orthonormality of the vectors, residuals of the leading pairs."""
    A = _device_cube(eng, torch, 2, 2, m, n)
    W = eng.gram(A, impl=1)
    r = W.shape[1]
    G = W.cpu().numpy().astype(np.complex128).transpose(0, 2, 1)
    eng.set_option("eig_impl", impl)
    try:
        lam, info = eng.eigh_jacobi(W)
    finally:
        eng.set_option("eig_impl", 0)
    lam, info, Wn = lam.cpu().numpy(), info.cpu().numpy(), W.cpu().numpy().astype(np.complex128)
    # info[:, 0]: Jacobi sweeps, or QL iterations (LAPACK's bound is 30 per eigenvalue)
    assert np.all(info[:, 1] == 1) and np.all(info[:, 0] <= (20 if impl == 1 else 30 * r))
    for b in range(G.shape[0]):
        ev = np.linalg.eigvalsh(G[b])[::-1]
        np.testing.assert_allclose(lam[b], ev, atol=3e-6 * ev[0])
        V = Wn[b] / np.linalg.norm(Wn[b], axis=1, keepdims=True)        # rows = eigenvectors
        assert np.abs(V.conj() @ V.T - np.eye(r)).max() < 1e-4
        top = np.argsort(-np.linalg.norm(Wn[b], axis=1))[:min(10, r)]
        for i in top:
            v = V[i]
            rq = np.real(v.conj() @ G[b] @ v)
            assert np.linalg.norm(G[b] @ v - rq * v) <= 2e-5 * ev[0]


def test_small_svd_stage(eng, torch):
    for (m, n) in [(64, 64), (16, 360), (360, 16), (33, 70), (1, 40), (40, 1), (2, 2)]:
        A = _device_cube(eng, torch, 2, 2, m, n)
        U, S, Vt, info = eng.svd_small(A)
        a = A.cpu().numpy()
        Uh, Sh, Vh, ih = U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy(), info.cpu().numpy()
        assert np.all(ih[:, 1] == 1), (m, n, ih)
        for b in range(a.shape[0]):
            parity.check_factors(a[b], Uh[b], Sh[b], Vh[b], min(m, n), label=f"small {m}x{n}")


# ---------------------------------------------------------------------------------------- edge cases
def test_empty_batch_and_bad_arguments(eng, torch):
    from visco_b200.compress_ms import apply_svd, apply_svd_batched
    from visco_b200.decompress_ms import reconstruct_vis, reconstruct_vis_batched
    assert apply_svd_batched(np.zeros((0, 8, 8), np.complex64)) == []
    assert reconstruct_vis_batched([]).shape[0] == 0
    with pytest.raises(ValueError):
        apply_svd(np.zeros((4, 4, 4), np.complex64))
    with pytest.raises(ValueError):
        reconstruct_vis(np.zeros((4, 2), np.complex64), np.zeros(3, np.float32), np.zeros((2, 5), np.complex64))
    A = torch.zeros((1, 8, 8), dtype=torch.complex64, device="cuda:0")
    with pytest.raises(ValueError):
        eng.compress(A, compressionrank=4, kmax=2)          # kmax too small


def test_non_finite_input_is_an_error(eng):
    from visco_b200.compress_ms import apply_svd
    a = np.ones((32, 48), np.complex64)
    a[3, 5] = np.nan
    with pytest.raises(ValueError):
        apply_svd(a, compressionrank=2)
    b = np.ones((128, 256), np.complex64)
    b[7, 9] = np.inf
    with pytest.raises(ValueError):
        apply_svd(b, compressionrank=2)


def test_zero_and_rank_deficient_matrices(eng):
    from visco_b200.compress_ms import apply_svd
    from visco_b200.decompress_ms import reconstruct_vis
    z = np.zeros((40, 56), np.complex64)
    U, S, Vt = apply_svd(z, compressionrank=3)
    assert S.shape == (3,) and not S.any()
    assert not reconstruct_vis(U, S, Vt).any()
    rng = np.random.default_rng(5)
    for (m, n) in [(48, 64), (150, 300)]:
        lo = (rng.standard_normal((m, 2)) + 1j * rng.standard_normal((m, 2))) @ \
             (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)))
        lo = lo.astype(np.complex64)
        U, S, Vt = apply_svd(lo, decorrelation=0.999)
        assert len(S) <= 2
        rec = reconstruct_vis(U, S, Vt)
        U2, S2, Vt2 = apply_svd(lo, compressionrank=2)
        rec2 = reconstruct_vis(U2, S2, Vt2)
        assert np.linalg.norm(lo - rec2) <= 1e-5 * np.linalg.norm(lo)
        assert np.linalg.norm(lo - rec) <= 0.05 * np.linalg.norm(lo)


def test_ragged_ranks_in_one_batch(eng):
    """decompression batches mix ranks (decompress_ms.py:188-199): padded factors + ranks."""
    from oracle import visco_oracle as vo
    from visco_b200.decompress_ms import reconstruct_vis_batched
    rng = np.random.default_rng(11)
    facs, refs = [], []
    for k in (1, 7, 3, 12):
        a = (rng.standard_normal((40, 72)) + 1j * rng.standard_normal((40, 72))).astype(np.complex64)
        u, s, vt = vo.ref_apply_svd(a, compressionrank=k)
        facs.append((u, s, vt))
        refs.append(vo.ref_reconstruct_vis(u, s, vt))
    out = reconstruct_vis_batched(facs)
    for b in range(4):
        np.testing.assert_allclose(out[b], refs[b], atol=3e-5 * np.abs(refs[b]).max())


# ---------------------------------------------------------------------------------------- size-independent properties
def test_properties_at_kat7_config_size(eng, torch):
    """BASELINE.json configs[1] at full size: 28 baselines x 4 corr x 256 x 1024, k = 8. No oracle at this size
    beyond a sample; properties: error^2 + retained energy == total energy (projection), idempotence of
    compress(reconstruct(.)), orthonormal factors."""
    nbl, ncorr, m, n, k = 28, 4, 256, 1024, 8
    A = _device_cube(eng, torch, nbl, ncorr, m, n)
    U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
    out = eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    assert int(ranks.min()) == k and int(ranks.max()) == k
    st = stats.cpu().numpy().astype(np.float64)
    assert np.all(st[:, 3] == 1)
    err2 = (A - out).abs().pow(2).sum(dim=(1, 2)).double().cpu().numpy()
    tot = A.abs().pow(2).sum(dim=(1, 2)).double().cpu().numpy()
    kept = S.double().pow(2).sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(err2 + kept, tot, rtol=2e-5)
    np.testing.assert_allclose(st[:, 0], tot, rtol=1e-4)
    I = torch.eye(k, dtype=torch.complex64, device="cuda:0")
    assert float((U.transpose(1, 2).conj() @ U - I).abs().max()) < 1e-4
    assert float((Vt @ Vt.transpose(1, 2).conj() - I).abs().max()) < 1e-4
    # idempotence: compressing the rank-8 reconstruction returns the same singular values and reconstruction
    U2, S2, Vt2, r2, _ = eng.compress(out, compressionrank=k)
    out2 = eng.reconstruct(U2, S2, Vt2, r2)
    torch.cuda.synchronize()
    assert float((S2 - S).abs().max() / S.max()) < 1e-5
    assert float((out2 - out).abs().max() / out.abs().max()) < 1e-4
    # oracle on a sample of the cube
    Ah = A[[0, 57, 111]].cpu().numpy()
    for i, b in enumerate((0, 57, 111)):
        parity.check_factors(Ah[i], U[b].cpu().numpy(), S[b].cpu().numpy(), Vt[b].cpu().numpy(), k,
                             compressionrank=k, label=f"kat7 b={b}")


# ---------------------------------------------------------------------------------------- fixed-rank fast path
@pytest.mark.parametrize("k", [1, 2, 4])
def test_fixed_rank_fast_path_matches_oracle_and_full_solver(eng, torch, k):
    """Blocked subspace iteration (compressionrank <= 4): signal-dominated matrices are solved by it, noise-dominated
    ones fall back to the full Jacobi solver inside the same call; both must satisfy the oracle tolerances and agree
    with the full solver."""
    A = _device_cube(eng, torch, 6, 4, 256, 1024, nbl_total=28, bl_offset=20)
    try:
        eng.set_option("topk", 1)
        Uf, Sf, Vf, rf, _ = eng.compress(A, compressionrank=k)
        eng.set_option("topk", 0)
        U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
        torch.cuda.synchronize()
    finally:
        eng.set_option("topk", 0)
    st = stats.cpu().numpy()
    assert np.all(st[:, 3] == 1)
    np.testing.assert_allclose(S.cpu().numpy(), Sf.cpu().numpy(), rtol=2e-6)
    a = A.cpu().numpy()
    Uh, Sh, Vh = U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy()
    for b in (0, 1, 3, 12, 13, 23):
        parity.check_factors(a[b], Uh[b], Sh[b], Vh[b], k, compressionrank=k, label=f"fast path k={k} b={b}")
        tot = float(np.sum(np.abs(a[b].astype(np.complex128)) ** 2))
        assert abs(st[b, 0] - tot) <= 1e-4 * tot                 # ||A||_F^2 comes from the Gram trace on this path


def test_fast_path_leaves_hard_spectra_to_the_full_solver(eng, torch):
    """Flat spectrum (pure noise) and an exactly rank-2 matrix with compressionrank=3: no gap at the cut, the iteration
    must give up (or certify) and the result must still be right."""
    from visco_b200.compress_ms import apply_svd
    rng = np.random.default_rng(3)
    noise = (rng.standard_normal((128, 512)) + 1j * rng.standard_normal((128, 512))).astype(np.complex64)
    U, S, Vt = apply_svd(noise, compressionrank=3)
    parity.check_factors(noise, U, S, Vt, 3, compressionrank=3, label="fast path on noise")
    lo = ((rng.standard_normal((160, 2)) + 1j * rng.standard_normal((160, 2))) @
          (rng.standard_normal((2, 400)) + 1j * rng.standard_normal((2, 400)))).astype(np.complex64)
    U, S, Vt = apply_svd(lo, compressionrank=3)
    assert S[2] <= 1e-3 * S[0]
    rec = (U * S) @ Vt
    assert np.linalg.norm(lo - rec) <= 1e-5 * np.linalg.norm(lo)


# ---------------------------------------------------------------------------------------- raw C ABI behaviour
def test_c_abi_status_codes_and_messages(eng, torch):
    """Error convention of the boundary: integer status + vk_last_error, no exceptions across the ABI."""
    import ctypes as C
    from visco_b200 import _lib
    lib = eng.lib
    A = torch.zeros((2, 8, 8), dtype=torch.complex64, device="cuda:0")
    U = torch.empty((2, 8, 8), dtype=torch.complex64, device="cuda:0")
    S = torch.empty((2, 8), dtype=torch.float32, device="cuda:0")
    Vt = torch.empty((2, 8, 8), dtype=torch.complex64, device="cuda:0")
    rk = torch.empty((2,), dtype=torch.int32, device="cuda:0")
    st = torch.empty((2, 4), dtype=torch.float32, device="cuda:0")
    p = lambda t: C.c_void_p(t.data_ptr())
    call = lambda *a: lib.vk_compress_batched(eng.h, *a)
    assert call(p(A), 2, 8, 8, 0, 0.0, 8, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_OK
    assert call(None, 2, 8, 8, 0, 0.0, 8, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_EINVAL
    assert b"null" in lib.vk_last_error(eng.h)
    assert call(p(A), 2, 0, 8, 0, 0.0, 8, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_EINVAL        # m = 0
    assert call(p(A), 2, 8, 8, 0, -0.5, 8, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_EINVAL      # decorrelation < 0
    assert call(p(A), 2, 8, 8, 4, 0.0, 2, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_EINVAL       # kmax < rank
    assert b"kmax" in lib.vk_last_error(eng.h)
    assert call(p(A), 2, 8, 8, 0, 0.0, 9, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_EINVAL       # kmax > min(m, n)
    assert call(p(A), 2, 8, 8, 0, 0.0, 8, p(U), p(S), p(Vt), p(rk), p(st), p(A), 16) == _lib.VK_EINVAL      # workspace too small
    assert call(p(A), 0, 8, 8, 0, 0.0, 8, p(U), p(S), p(Vt), p(rk), p(st), None, 0) == _lib.VK_OK           # empty batch
    assert lib.vk_set_option(eng.h, b"no_such_option", 1.0) == _lib.VK_EINVAL
    assert lib.vk_reconstruct_batched(eng.h, p(U), p(S), p(Vt), None, 2, 8, 8, 0, p(A)) == _lib.VK_EINVAL   # kmax = 0
    assert lib.vk_workspace_bytes(eng.h, 64, 512, 4096, 512) >= 64 * 512 * 512 * 8
    big = torch.empty((1, 4, 4), dtype=torch.complex64, device="cuda:0")
    assert lib.vk_svd_jacobi_small_batched(eng.h, p(big), 1, 4096, 4096, p(U), p(S), p(Vt), None) == _lib.VK_EINVAL


def test_handles_are_independent_and_thread_safe(torch):
    """One handle per host thread (SURVEY 8b threading): two engines on the same GPU used concurrently give the same
    answer as a serial run; destroying one leaves the other usable."""
    import threading
    from visco_b200.engine import Engine, get_engine
    base = get_engine(0)
    A = _device_cube(base, torch, 4, 4, 128, 256)
    ref = base.compress(A, compressionrank=5)[1].cpu().numpy()
    engines = [Engine(0), Engine(0)]
    out = [None, None]

    def work(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                out[i] = engines[i].compress(A, compressionrank=5)[1]
            s.synchronize()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(2):
        np.testing.assert_allclose(out[i].cpu().numpy(), ref, rtol=1e-6)
    engines[0].close()
    np.testing.assert_allclose(engines[1].compress(A, compressionrank=5)[1].cpu().numpy(), ref, rtol=1e-6)
    engines[1].close()


def test_internal_scheduling_options_do_not_change_results(eng, torch):
    """chunking, stream groups, the generic Jacobi kernels and the choice of eigensolver are performance knobs only."""
    A = _device_cube(eng, torch, 3, 4, 192, 384)
    keys = ("chunk", "jacobi_groups", "jacobi_generic", "gram_impl", "gemm_impl", "eig_impl")
    try:
        for eig in (1, 2):
            eng.set_option("eig_impl", eig)
            base = eng.compress(A, decorrelation=0.97)
            if eig == 1:
                jacobi = base
            variants = ({"chunk": 5}, {"gram_impl": 1}, {"gemm_impl": 1})
            if eig == 1:
                variants += ({"jacobi_groups": 1}, {"jacobi_groups": 4}, {"jacobi_generic": 1})
            for opts in variants:
                for k_, v_ in opts.items():
                    eng.set_option(k_, v_)
                got = eng.compress(A, decorrelation=0.97)
                for k_ in opts:
                    eng.set_option(k_, 0)
                assert torch.equal(got[3], base[3]), (eig, opts)
                assert float(((got[1] - base[1]).abs() / base[1].clamp_min(1e-20)).max()) < 2e-5, (eig, opts)
        # the two eigensolvers against each other: same ranks, same singular values
        assert torch.equal(base[3], jacobi[3])
        assert float(((base[1] - jacobi[1]).abs() / jacobi[1].clamp_min(1e-20)).max()) < 1e-4
    finally:
        for k_ in keys:
            eng.set_option(k_, 0)


@pytest.mark.parametrize("k", [3, 4])
def test_multiple_and_clustered_leading_singular_values(eng, k):
    """Fixed rank on the Gram path with exactly multiple, nearly multiple and vanishing leading singular values: the
    leading-eigenpair path (bisection + twisted factorisation) has to hand true clusters to the full solver and still
    return an orthonormal basis of the optimal subspace."""
    from visco_b200.compress_ms import apply_svd
    from visco_b200.decompress_ms import reconstruct_vis
    rng = np.random.default_rng(3)
    m, n = 128, 256
    Q1, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    Q2, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    cases = [np.ones(m),
             np.concatenate([[5, 5, 5, 2, 1], 0.01 * np.ones(m - 5)]),
             np.concatenate([[5, 5 * (1 - 1e-6), 3, 2, 1], 0.01 * rng.random(m - 5)]),
             np.concatenate([[5, 5 * (1 - 1e-4), 3, 2, 1], 0.01 * rng.random(m - 5)]),
             np.concatenate([[3, 1], np.zeros(m - 2)])]
    for sv in cases:
        A = ((Q1 * sv[None, :]) @ Q2[:m]).astype(np.complex64)
        U, S, Vt = apply_svd(A, compressionrank=k)
        s_sorted = np.sort(sv)[::-1]
        np.testing.assert_allclose(S, s_sorted[:k], rtol=1e-4, atol=2e-6 * s_sorted[0])
        rec = reconstruct_vis(U, S, Vt)
        optimal = np.sqrt(np.sum(s_sorted[k:] ** 2))
        err = np.linalg.norm(A.astype(np.complex128) - rec)
        assert abs(err - optimal) <= 1e-5 * optimal + 5e-6 * np.linalg.norm(A), (sv[:5], err, optimal)
        assert np.abs(U.conj().T @ U - np.eye(k)).max() < 1e-4


@pytest.mark.parametrize("B,m,n,kw", [
    (2, 1024, 1100, dict(compressionrank=16)),   # largest size of the direct eigensolver, leading-pair path (16 vectors)
    (2, 1024, 1100, dict(decorrelation=0.9)),    # ... full QL path
    (2, 700, 2000, dict(compressionrank=40)),    # fixed rank above the leading-pair limit: full path
    (1, 1100, 1200, dict(compressionrank=8)),    # min(m, n) > 1024: cyclic Jacobi solver
    (2, 1500, 130, dict(decorrelation=0.95)),    # tall, Gram on the column side
])
def test_gram_path_size_limits_of_the_eigensolvers(eng, torch, B, m, n, kw):
    A = _device_cube(eng, torch, B, 1, m, n)
    U, S, Vt, ranks, stats = eng.compress(A, **kw)
    torch.cuda.synchronize()
    Ah, Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (A, U, S, Vt, ranks, stats))
    assert np.all(st[:, 3] == 1)
    for b in range(B):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"{m}x{n} {kw}", **kw)


def test_direct_solver_hands_unconverged_passes_to_jacobi(eng, torch):
    """The QL iteration is limited per eigenvalue like LAPACK's; a matrix that hits the limit sends its pass back
    through the cyclic Jacobi solver. Forced here with a limit of one iteration."""
    A = _device_cube(eng, torch, 2, 4, 160, 320)
    try:
        eng.set_option("eigvec_impl", 1)              # the QL route (the default full-spectrum route does not iterate)
        base = eng.compress(A)                        # full rank: every matrix takes the QL iteration
        assert float(base[4][:, 2].min()) > 100      # QL iterations: the direct solver ran
        eng.set_option("ql_maxit", 1)
        got = eng.compress(A)
    finally:
        eng.set_option("ql_maxit", 60)
        eng.set_option("eigvec_impl", 0)
    torch.cuda.synchronize()
    assert float(got[4][:, 3].min()) == 1 and float(got[4][:, 2].max()) <= 30   # converged, in Jacobi sweeps
    assert torch.equal(got[3], base[3])
    assert float(((got[1] - base[1]).abs() / base[1].clamp_min(1e-20)).max()) < 1e-4


def test_ill_conditioned_matrices_are_redone_without_a_gram_product(eng, torch):
    """A float32 Gram matrix resolves eigenvalues only down to ~1e-7 of the largest one: singular values a few 1e-3 below
    sigma_1 would come out wrong (and their vectors mixed). Matrices whose RETAINED values reach that far are detected and
    done again by one-sided Jacobi on the matrix itself; the result has to meet the same tolerances as everywhere."""
    rng = np.random.default_rng(42)
    m, n = 150, 268
    Q1, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    Q2, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    sv = np.sort(np.concatenate([100.0 * (1 + rng.random(5)), 0.05 * (0.5 + rng.random(m - 5))]))[::-1]
    A = np.stack([((Q1 * sv[None, :]) @ Q2[:m]).astype(np.complex64),
                  (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))).astype(np.complex64)])
    Ad = torch.from_numpy(A).cuda()
    for kw in (dict(compressionrank=24), dict(), dict(decorrelation=0.9999999)):
        U, S, Vt, ranks, stats = eng.compress(Ad, **kw)
        torch.cuda.synchronize()
        Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (U, S, Vt, ranks, stats))
        assert np.all(st[:, 3] == 1)
        for b in range(2):
            k = int(rk[b])
            parity.check_factors(A[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"ill-conditioned {kw} b={b}", **kw)
    # the well-conditioned matrix of the batch and a truncation that stops above the gap do not take the detour
    eng.set_option("illcond_thr", 0.0)
    try:
        U0, S0, Vt0, r0, _ = eng.compress(Ad, compressionrank=5)
    finally:
        eng.set_option("illcond_thr", 0.005)
    U1, S1, Vt1, r1, _ = eng.compress(Ad, compressionrank=5)
    # (not bit-identical from call to call: the refined singular values are accumulated with atomics)
    assert torch.equal(r0, r1) and float(((S0 - S1).abs() / S0.clamp_min(1e-20)).max()) < 2e-6


def test_odd_row_count_with_large_rank(eng, torch):
    """m odd, kmax > 16 and n % 16 == 0: the tcgen05 V-formation cannot take an odd m (16-byte tensor-map strides), the
    SIMT GEMM has to (found by tools/stress_parity.py: the call used to fail with a tensor-map error)."""
    for (m, n, k) in [(147, 528, 18), (137, 208, 40)]:
        A = _device_cube(eng, torch, 1, 2, m, n)
        U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
        out = eng.reconstruct(U, S, Vt, ranks)
        torch.cuda.synchronize()
        Ah, Uh, Sh, Vh, oh = (x.cpu().numpy() for x in (A, U, S, Vt, out))
        for b in range(2):
            parity.check_factors(Ah[b], Uh[b], Sh[b], Vh[b], k, compressionrank=k, label=f"{m}x{n} k={k}")
            parity.check_reconstruction(Uh[b], Sh[b], Vh[b], oh[b], label=f"{m}x{n} k={k}")


def test_energy_rule_through_the_leading_pair_path(eng, torch):
    """Option topk = 2: with the energy rule the rank is estimated on the device from all eigenvalues (bisection) and
    matrices that end up with a small rank take the leading-eigenpair path with their own k. Same ranks and factors as
    the full QL path; matrices with a large rank (the noise-dominated cross hands) keep the full path."""
    A = _device_cube(eng, torch, 6, 4, 192, 640)
    base = eng.compress(A, decorrelation=0.9)
    try:
        eng.set_option("topk", 2)
        got = eng.compress(A, decorrelation=0.9)
    finally:
        eng.set_option("topk", 0)
    torch.cuda.synchronize()
    st = got[4].cpu().numpy()
    assert np.all(st[:, 3] == 1)
    assert (st[:, 2] == 0).sum() >= 6 and (st[:, 2] > 0).sum() >= 6      # both routes were taken
    assert torch.equal(got[3], base[3])
    Ah, Uh, Sh, Vh, rk = (x.cpu().numpy() for x in (A, got[0], got[1], got[2], got[3]))
    for b in range(A.shape[0]):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, decorrelation=0.9, label=f"energy topk=2 b={b}")


def test_stage_and_kernel_timers_of_the_c_abi(eng, torch):
    """vk_last_stage_ms / vk_last_eig_ms (option stage_timing = 1): the slots add up and name the kernels that ran."""
    A = _device_cube(eng, torch, 4, 4, 256, 512)
    try:
        eng.set_option("stage_timing", 1)
        eng.compress(A, compressionrank=4)
        st, eig = eng.last_stage_ms(), eng.last_eig_ms()
        assert st["gram"] > 0 and st["jacobi"] > 0 and st["factors"] > 0 and st["total"] >= st["jacobi"]
        assert eig["tridiag"] > 0 and eig["leading_pairs"] > 0               # fixed small rank: leading eigenpairs
        assert abs(sum(eig.values()) - st["jacobi"]) <= 0.2 * st["jacobi"] + 0.05
        eng.compress(A, decorrelation=0.99)
        eig = eng.last_eig_ms()
        assert eig["ql"] > 0 and eig["reflectors"] > 0 and eig["rotations"] > 0   # energy rule: the full QL route
    finally:
        eng.set_option("stage_timing", 0)


@pytest.mark.parametrize("scale", [1e25, 3e-26])
@pytest.mark.parametrize("shape,kw", [((256, 512), dict(compressionrank=6)), ((96, 300), dict(decorrelation=0.97)),
                                       ((200, 16), dict(compressionrank=4))])
def test_inputs_whose_squares_leave_the_float32_range(eng, torch, scale, shape, kw):
    """|a| ~ 1e25 overflows, |a| ~ 1e-26 vanishes in a float32 Gram product (LAPACK's cgesdd scales its input). Such matrices
    are detected by their trace and done again without a Gram product, on data scaled by a power of two; a genuinely
    non-finite input is still an error."""
    rng = np.random.default_rng(5)
    m, n = shape
    B = 3
    a = (rng.standard_normal((B, m, n)) + 1j * rng.standard_normal((B, m, n))).astype(np.complex64)
    a[:, : m // 4] *= 6.0                                        # some structure in the spectrum
    a[1] *= np.float32(scale)                                    # one matrix of the batch is out of range, its neighbours are not
    U, S, Vt, ranks, stats = (x.cpu().numpy() for x in eng.compress(torch.from_numpy(a).cuda(), **kw))
    for b in range(B):
        k = int(ranks[b])
        assert np.isfinite(S[b, :k]).all() and S[b, 0] > 0
        parity.check_factors(a[b], U[b, :, :k], S[b, :k], Vt[b, :k], k, label=f"scaled {scale} {shape} b={b}", **kw)
    bad = a.copy()
    bad[2, 1, 1] = np.nan
    with pytest.raises(ValueError):
        eng.compress(torch.from_numpy(bad).cuda(), **kw)
