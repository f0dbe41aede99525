"""GPU parity tests at the BASELINE.json configurations the round-1 suite did not reach:
C3 (512 x 4096, decorrelation 0.99), C4 (a >= 512-matrix batch of 64 x 64), C5 (128 x 2048 reconstruction for EVERY
k = 1..32 against the oracle), plus the padding contract of vk_reconstruct_batched (ranks < kmax with garbage beyond).
Every call goes through the C ABI; the oracle is only the checker."""
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


def _device_cube(eng, torch, nbl, ncorr, m, n, nbl_total=None, bl_offset=0):
    A = torch.empty((nbl * ncorr, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, ncorr, bl_offset=bl_offset, nbl_total=nbl_total or nbl)
    return A


# ------------------------------------------------------------------------------------------------ C3
def test_c3_meerkat_shape_energy_rule(eng, torch):
    """BASELINE configs[2] shape: 512 x 4096, decorrelation 0.99 (energy fraction 0.9801), a short and a long baseline,
    parallel and cross hands; full check_factors against the oracle (reference compress_ms.py:295-363)."""
    m, n, dec = 512, 4096, 0.99
    for off in (40, 1900):                       # fringe-rate scale 30 (b + 1) / 2080: short and long baseline
        A = _device_cube(eng, torch, 1, 4, m, n, nbl_total=2080, bl_offset=off)
        U, S, Vt, ranks, stats = eng.compress(A, decorrelation=dec)
        out = eng.reconstruct(U, S, Vt, ranks)
        torch.cuda.synchronize()
        rk, st = ranks.cpu().numpy(), stats.cpu().numpy()
        assert np.all(st[:, 3] == 1)
        for b in (0, 1):                         # XX (signal + noise) and XY (noise dominated, near-full rank)
            k = int(rk[b])
            a = A[b].cpu().numpy()
            Uh, Sh, Vh = U[b, :, :k].cpu().numpy(), S[b, :k].cpu().numpy(), Vt[b, :k].cpu().numpy()
            res = parity.check_factors(a, Uh, Sh, Vh, k, decorrelation=dec, label=f"C3 512x4096 dec0.99 bl{off} b={b}")
            assert 1 <= res["k"] <= m
            parity.check_reconstruction(Uh, Sh, Vh, out[b].cpu().numpy(), label=f"C3 recon bl{off} b={b}")
            assert not U[b, :, k:].any().item() and not Vt[b, k:].any().item()


def test_c3_shape_fixed_rank(eng, torch):
    A = _device_cube(eng, torch, 1, 4, 512, 4096, nbl_total=2080, bl_offset=700)
    U, S, Vt, ranks, stats = eng.compress(A, compressionrank=6)
    torch.cuda.synchronize()
    for b in (0, 2):
        parity.check_factors(A[b].cpu().numpy(), U[b].cpu().numpy(), S[b].cpu().numpy(), Vt[b].cpu().numpy(), 6,
                             compressionrank=6, label=f"C3 512x4096 k6 b={b}")


# ------------------------------------------------------------------------------------------------ C4
@pytest.mark.parametrize("small_impl", [1, 0])
@pytest.mark.parametrize("kw", [dict(compressionrank=8), dict(decorrelation=0.95)])
def test_c4_small_matrix_batch(eng, torch, kw, small_impl):
    """BASELINE configs[3]: 64 x 64 matrices, a 520-matrix batch (130 baselines x 4), every matrix checked against the
    oracle - through the one-sided Jacobi path the configuration names ("small_impl" = 1) and through the default route
    (Gram product + warp-level tridiagonalisation, tridiag_small.cu)."""
    nbl, ncorr, m, n = 130, 4, 64, 64
    assert not eng.uses_small_path(m, n) and eng.uses_small_path(32, 64)
    A = _device_cube(eng, torch, nbl, ncorr, m, n, nbl_total=2080, bl_offset=975)
    eng.set_option("small_impl", small_impl)
    try:
        U, S, Vt, ranks, stats = eng.compress(A, **kw)
    finally:
        eng.set_option("small_impl", 0)
    out = eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    Ah, Uh, Sh, Vh, rk, st, oh = (x.cpu().numpy() for x in (A, U, S, Vt, ranks, stats, out))
    assert np.all(st[:, 3] == 1)
    for b in range(nbl * ncorr):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"C4 64x64 {kw} b={b}", **kw)
        if b % 16 == 0:
            parity.check_reconstruction(Uh[b, :, :k], Sh[b, :k], Vh[b, :k], oh[b], label=f"C4 recon {kw} b={b}")


# ------------------------------------------------------------------------------------------------ C5
def _random_factors(rng, B, m, n, k):
    U = ((rng.standard_normal((B, m, k)) + 1j * rng.standard_normal((B, m, k))) / np.sqrt(2 * m)).astype(np.complex64)
    Vt = ((rng.standard_normal((B, k, n)) + 1j * rng.standard_normal((B, k, n))) / np.sqrt(2 * n)).astype(np.complex64)
    S = (100.0 * np.exp(-0.25 * np.arange(k))[None, :] * (1 + 0.1 * rng.random((B, k)))).astype(np.float32)
    return U, S, Vt


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_c5_reconstruction_every_rank(eng, torch, k):
    """BASELINE configs[4]: 128 x 2048 reconstruction from given factors for every k = 1..32, against
    oracle.ref_reconstruct_vis (reference decompress_ms.py:107-131) - not against a library matmul."""
    from oracle import visco_oracle as vo
    rng = np.random.default_rng(1000 + k)
    B, m, n = 6, 128, 2048
    U, S, Vt = _random_factors(rng, B, m, n, k)
    out = eng.reconstruct(torch.from_numpy(U).cuda(), torch.from_numpy(S).cuda(), torch.from_numpy(Vt).cuda()).cpu().numpy()
    for b in range(B):
        parity.check_reconstruction(U[b], S[b], Vt[b], out[b], label=f"C5 128x2048 k{k}")
        ref = vo.ref_reconstruct_vis(U[b], S[b], Vt[b])
        # Frobenius-level agreement with the reference's own complex64 product
        assert np.linalg.norm(out[b] - ref) <= 3e-6 * np.linalg.norm(ref), (k, b)


@pytest.mark.parametrize("kmax,n", [(8, 256), (16, 256), (32, 256), (32, 250), (5, 33), (40, 512)])
def test_reconstruct_uses_only_the_first_ranks_modes(eng, torch, kmax, n):
    """include/visco_b200.h: 'using the first ranks[b] modes'. Whatever sits beyond ranks[b] (NaN, Inf, stale data from
    torch.empty) must not reach the result on ANY of the three kernels (small-k streaming, tcgen05 GEMM, SIMT GEMM)."""
    from oracle import visco_oracle as vo
    rng = np.random.default_rng(7 + kmax + n)
    B, m = 5, 96
    U, S, Vt = _random_factors(rng, B, m, n, kmax)
    ranks = np.array([kmax, max(1, kmax // 2), 1, kmax - 1, 0], np.int32)
    Ug, Sg, Vg = U.copy(), S.copy(), Vt.copy()
    for b in range(B):
        kb = int(ranks[b])
        Ug[b, :, kb:] = np.nan if b % 2 == 0 else 1e30
        Sg[b, kb:] = np.inf if b % 2 == 0 else 7.0
        Vg[b, kb:] = np.nan if b % 2 == 1 else -3.0
    out = eng.reconstruct(torch.from_numpy(Ug).cuda(), torch.from_numpy(Sg).cuda(), torch.from_numpy(Vg).cuda(),
                          torch.from_numpy(ranks).cuda()).cpu().numpy()
    assert np.isfinite(out.view(np.float32)).all()
    for b in range(B):
        kb = int(ranks[b])
        ref = vo.ref_reconstruct_vis(U[b, :, :kb], S[b, :kb], Vt[b, :kb]) if kb else np.zeros((m, n), np.complex64)
        scale = max(float(np.abs(ref).max()), 1e-6)
        assert np.abs(out[b] - ref).max() <= 2e-5 * scale * max(1.0, np.sqrt(kb)), (kmax, n, b)


# ------------------------------------------------------------------------------------------------ layout validation
def test_layout_indices_are_validated(eng, torch):
    """A correlation plane or a row outside the column raises ValueError (numpy indexing in the reference raises
    IndexError, decompress_ms.py:216-232) instead of writing outside the buffer."""
    data = torch.zeros((10, 4, 2), dtype=torch.complex64, device="cuda:0")
    cube = torch.ones((1, 5, 4), dtype=torch.complex64, device="cuda:0")
    rows = torch.arange(5, dtype=torch.int32, device="cuda:0").reshape(1, 5)
    with pytest.raises(ValueError):
        eng.scatter_baselines(cube, data, rows, torch.tensor([[3]], dtype=torch.int32, device="cuda:0"))
    with pytest.raises(ValueError):
        eng.scatter_baselines(cube, data, rows + 8, torch.tensor([[1]], dtype=torch.int32, device="cuda:0"))
    with pytest.raises(ValueError):
        eng.gather_baselines(data, rows, torch.tensor([[-1]], dtype=torch.int32, device="cuda:0"))
    assert not data.any().item()
    eng.scatter_baselines(cube, data, rows, torch.tensor([[1]], dtype=torch.int32, device="cuda:0"))
    assert data[:5, :, 1].eq(1).all().item() and not data[:, :, 0].any().item() and not data[5:].any().item()


def test_two_correlation_store_round_trip(eng, torch, tmp_path):
    """CORR_TYPE [XX, YY] column (ncorr = 2) through compress_visdata / construct_main_ds with --correlation-optimized
    and per-correlation leaves: the decompressor maps leaf names to planes through the stored POLARIZATION/CORR_TYPE,
    as the compressor does (reference compress_ms.py:601-602, 662)."""
    from visco_b200.compress_ms import compress_full_ms
    from visco_b200.decompress_ms import construct_main_ds
    from visco_b200.msdata import VisData
    rng = np.random.default_rng(3)
    nant, ntime, nchan = 4, 24, 16
    a1, a2 = np.array([(i, j) for i in range(nant) for j in range(i + 1, nant)]).T
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    nrow = ant1.size
    t = np.arange(nrow)[:, None, None] / nrow
    data = (np.exp(2j * np.pi * (3 * t + 0.1 * np.arange(nchan)[None, :, None] / nchan))
            * np.array([1.0, 2.0])[None, None, :]).astype(np.complex64)
    data += 1e-3 * (rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape)).astype(np.complex64)
    vis = VisData(data=data, antenna1=ant1, antenna2=ant2, antenna_names=[f"A{i}" for i in range(nant)], corr_types=[9, 12])
    src = str(tmp_path / "two_corr.npz")
    vis.save(src)
    for opt in (True, False):
        store = str(tmp_path / f"store_{int(opt)}.zarr")
        compress_full_ms(src, store, correlation="XX,YY", correlation_optimized=opt, compressionrank=3)
        back = construct_main_ds(store, "COMPRESSED_DATA", 50)
        assert back.data.shape == data.shape and back.corr_types == [9, 12]
        err = np.linalg.norm(back.data - data) / np.linalg.norm(data)
        assert err < 5e-3, (opt, err)


# ------------------------------------------------------------------------------------------------ eigenvector paths
@pytest.mark.parametrize("shape,dec", [((256, 1024), 0.98), ((128, 512), None), ((512, 2048), 0.99)])
def test_full_spectrum_paths_agree(eng, torch, shape, dec):
    """The twisted-factorisation + Newton-Schulz + GEMM eigenvector path ("eigvec_impl" = 0, default) against the implicit
    QL path ("eigvec_impl" = 1) on the same cube: same ranks, singular values to 2e-5, both within the parity bounds."""
    m, n = shape
    A = _device_cube(eng, torch, 2, 4, m, n, nbl_total=64, bl_offset=20)
    res = {}
    for impl in (1, 0):
        eng.set_option("eigvec_impl", impl)
        try:
            U, S, Vt, ranks, stats = eng.compress(A, decorrelation=dec)
            torch.cuda.synchronize()
            res[impl] = tuple(x.cpu().numpy() for x in (U, S, Vt, ranks, stats))
        finally:
            eng.set_option("eigvec_impl", 0)
    assert np.all(res[0][4][:, 3] == 1) and np.all(res[1][4][:, 3] == 1)
    Ah = A.cpu().numpy()
    for b in range(A.shape[0]):
        k0, k1 = int(res[0][3][b]), int(res[1][3][b])
        assert abs(k0 - k1) <= 1, (b, k0, k1)
        kk = min(k0, k1)
        np.testing.assert_allclose(res[0][1][b, :kk], res[1][1][b, :kk], rtol=2e-5, atol=2e-6 * res[1][1][b, 0])
        if b in (0, 1, 5):
            parity.check_factors(Ah[b], res[0][0][b, :, :k0], res[0][1][b, :k0], res[0][2][b, :k0], k0, decorrelation=dec,
                                 label=f"eigvec {shape} dec{dec} b={b}")


# ------------------------------------------------------------------------------------------------ round-2 kernels
@pytest.mark.parametrize("B,m,n,kw,opts", [
    (304, 200, 600, dict(compressionrank=5), {}),                # tridiag_sym.cu: ragged r, more matrices than SMs (two per SM)
    (12, 200, 600, dict(decorrelation=0.97), {}),                # ... ragged r, one matrix per SM, full spectrum through QL (r % 16)
    (8, 384, 1000, dict(decorrelation=0.98), {}),                # ... r = 384 (12 column chunks), full-spectrum path
    (6, 144, 400, dict(compressionrank=12), {}),                 # ... smallest sizes of that kernel
    (6, 512, 1024, dict(compressionrank=3), {"tridiag_variant": 2}),   # two-per-SM launch shape forced on a small batch
    (6, 256, 1024, dict(compressionrank=8), {"tridiag_impl": 1}),      # round 1's undeferred kernels stay selectable
    (10, 16, 4096, dict(compressionrank=4), {}),                 # tridiag_small.cu, one warp per matrix (r <= 33, too long for the Jacobi path)
    (10, 33, 3000, dict(decorrelation=0.9), {}),                 # ... odd r, full path
    (10, 64, 4096, dict(compressionrank=8), {}),                 # ... two warps per matrix
    (10, 50, 60, dict(decorrelation=0.95), {}),                  # 32 < r <= 64, Gram route by default
])
def test_round2_tridiagonalisation_kernels(eng, torch, B, m, n, kw, opts):
    A = _device_cube(eng, torch, B, 1, m, n)
    for k_, v_ in opts.items():
        eng.set_option(k_, v_)
    try:
        U, S, Vt, ranks, stats = eng.compress(A, **kw)
    finally:
        for k_ in opts:
            eng.set_option(k_, 0)
    torch.cuda.synchronize()
    Ah, Uh, Sh, Vh, rk, st = (x.cpu().numpy() for x in (A, U, S, Vt, ranks, stats))
    assert np.all(st[:, 3] == 1)
    for b in sorted(set(range(0, B, max(1, B // 6))) | {B - 1}):
        k = int(rk[b])
        parity.check_factors(Ah[b], Uh[b, :, :k], Sh[b, :k], Vh[b, :k], k, label=f"r2 {m}x{n} {kw} {opts} b={b}", **kw)


@pytest.mark.parametrize("m,n,k", [(256, 1024, 8), (200, 1000, 3), (128, 2048, 16), (96, 250, 5), (300, 2050, 7), (256, 1024, 1)])
def test_fused_factor_formation_matches_the_separate_kernels(eng, torch, m, n, k):
    """factors_fused_kernel (one cluster launch: U, refined sigma, normalised Vt, retained energy) against the five-launch
    path it replaces ("factors_impl" = 1) and against the oracle; ragged last strip (n % 256 != 0), k below the template
    width, n beyond eight strips (falls back to the separate kernels)."""
    A = _device_cube(eng, torch, 9, 1, m, n)
    res = {}
    for impl in (1, 0):
        eng.set_option("factors_impl", impl)
        try:
            res[impl] = [x.clone() for x in eng.compress(A, compressionrank=k)]
        finally:
            eng.set_option("factors_impl", 0)
    torch.cuda.synchronize()
    U0, S0, V0, r0, t0 = (x.cpu().numpy() for x in res[0])
    U1, S1, V1, r1, t1 = (x.cpu().numpy() for x in res[1])
    np.testing.assert_array_equal(r0, r1)
    np.testing.assert_array_equal(U0, U1)                          # same gather
    np.testing.assert_allclose(S0, S1, rtol=2e-6)                 # summation order differs (atomics vs fixed order)
    np.testing.assert_allclose(t0[:, 1], t1[:, 1], rtol=1e-5)
    assert np.abs(V0 - V1).max() <= 2e-6 * np.abs(V1).max()
    Ah = A.cpu().numpy()
    for b in (0, 8):
        parity.check_factors(Ah[b], U0[b], S0[b], V0[b], k, compressionrank=k, label=f"fused factors {m}x{n} k{k} b={b}")


@pytest.mark.parametrize("m,n,kw", [(64, 64, dict(compressionrank=8)), (64, 64, dict(decorrelation=0.95)), (48, 100, dict(decorrelation=0.9)),
                                    (100, 48, dict(compressionrank=5)), (33, 200, dict(compressionrank=4)), (130, 40, dict())])
def test_fused_small_gram_matches_the_separate_kernels(eng, torch, m, n, kw):
    """gram_small_kernel (Gram product + trace normalisation of a matrix with min(m, n) <= 64 in one CTA, both sides:
    rows for m <= n, columns for m > n; ragged extents, a contraction longer than one 64-step chunk) against the SIMT GEMM
    + normalisation pass it replaces ("gram_small" = 1) and against the oracle."""
    A = _device_cube(eng, torch, 10, 4, m, n)
    res = {}
    for impl in (1, 0):
        eng.set_option("gram_small", impl)
        try:
            res[impl] = [x.clone() for x in eng.compress(A, **kw)]
        finally:
            eng.set_option("gram_small", 0)
    torch.cuda.synchronize()
    assert torch.equal(res[0][3], res[1][3])                                     # ranks
    S0, S1 = res[0][1], res[1][1]
    assert float(((S0 - S1).abs() / S1.abs().clamp_min(1e-20)).max()) < 5e-6     # summation order differs
    Ah = A.cpu().numpy()
    U, S, Vt, ranks, stats = (x.cpu().numpy() for x in res[0])
    assert np.all(stats[:, 3] == 1)
    for b in (0, 1, 17, 39):
        k = int(ranks[b])
        parity.check_factors(Ah[b], U[b, :, :k], S[b, :k], Vt[b, :k], k, label=f"fused small gram {m}x{n} {kw} b={b}", **kw)


@pytest.mark.parametrize("m,n,k", [(64, 64, 8), (48, 70, 3), (100, 40, 15), (128, 128, 1)])
def test_packed_bisection_of_many_small_problems(eng, torch, m, n, k):
    """bisect_packed_kernel (batches of at least four matrices per SM with min(m, n) <= 128 and a fixed rank <= 15: several
    matrices per warp, one lane per eigenvalue) against the CTA-per-matrix kernel ("bisect_impl" = 1) and the oracle; the
    batch is not a whole number of CTAs."""
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    nbl = nsm + 3
    A = _device_cube(eng, torch, nbl, 4, m, n)
    res = {}
    for impl in (1, 0):
        eng.set_option("bisect_impl", impl)
        try:
            res[impl] = [x.clone() for x in eng.compress(A, compressionrank=k)]
        finally:
            eng.set_option("bisect_impl", 0)
    torch.cuda.synchronize()
    assert torch.equal(res[0][3], res[1][3])
    S0, S1 = res[0][1], res[1][1]
    assert float(((S0 - S1).abs() / S1.abs().clamp_min(1e-20)).max()) < 5e-6
    Ah = A.cpu().numpy()
    U, S, Vt, ranks, stats = (x.cpu().numpy() for x in res[0])
    assert np.all(stats[:, 3] == 1)
    B = A.shape[0]
    for b in (0, 1, 2, 3, B // 2, B - 2, B - 1):
        kk = int(ranks[b])
        parity.check_factors(Ah[b], U[b, :, :kk], S[b, :kk], Vt[b, :kk], kk, compressionrank=k, label=f"packed bisection {m}x{n} k{k} b={b}")


def test_remainder_split_of_the_eigensolver_changes_nothing(eng, torch):
    """More matrices than SMs with a small remainder: the remainder runs as its own sub-batch on a second stream
    (tridiag.cu, "tail_split"). The split may put the main sub-batch on the other launch shape of the tridiagonalisation
    (tile height and deferral depth differ), so the two runs agree to rounding, not bit for bit; both are checked against
    the oracle."""
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    B, m, n = nsm + 5, 160, 512
    A = _device_cube(eng, torch, B, 1, m, n)
    res = {}
    for off in (1, 0):
        eng.set_option("tail_split", off)
        try:
            res[off] = [x.clone() for x in eng.compress(A, decorrelation=0.97)]
        finally:
            eng.set_option("tail_split", 0)
    torch.cuda.synchronize()
    assert torch.equal(res[0][3], res[1][3])                                     # ranks
    assert float(((res[0][1] - res[1][1]).abs() / res[1][1].abs().clamp_min(1e-20)).max()) < 2e-5
    Ah = A.cpu().numpy()
    for off in (0, 1):
        U, S, Vt, ranks, stats = (x.cpu().numpy() for x in res[off])
        assert np.all(stats[:, 3] == 1)
        for b in (0, nsm - 1, nsm, B - 1):
            k = int(ranks[b])
            parity.check_factors(Ah[b], U[b, :, :k], S[b, :k], Vt[b, :k], k, decorrelation=0.97, label=f"tail split {off} b={b}")
