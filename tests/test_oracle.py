"""CPU: the oracle (oracle/visco_oracle.py) against fixtures produced by the reference's own function bodies
(tests/golden/make_golden.py), including matrices from the reference's sample Measurement Set."""
import numpy as np
import pytest

from oracle import visco_oracle as vo


def test_golden_has_sample_ms_and_synthetic(golden_cases):
    assert any(c.startswith("ms_") for c in golden_cases)
    assert any(c.startswith("syn_") for c in golden_cases)
    assert len(golden_cases) >= 12


def test_singular_values_match_reference(golden, golden_cases):
    for c in golden_cases:
        a = golden[f"{c}/A"]
        u, s, vt = vo.ref_apply_svd(a)
        assert s.dtype == np.float32 and u.dtype == np.complex64 and vt.dtype == np.complex64
        assert u.shape == (a.shape[0], min(a.shape)) and vt.shape == (min(a.shape), a.shape[1])
        np.testing.assert_allclose(s, golden[f"{c}/S"], rtol=1e-6, atol=1e-6 * float(golden[f"{c}/S"][0]))


def test_energy_rule_matches_reference(golden, golden_cases):
    decs = golden["decs"]
    for c in golden_cases:
        s = golden[f"{c}/S"]
        got = [vo.ref_find_n_decorrelation(s, float(d)) for d in decs]
        assert got == list(golden[f"{c}/n_dec"]), c


def test_energy_rule_uses_squared_decorrelation_and_float32():
    s = np.array([3.0, 2.0, 1.0, 0.5], np.float32)           # energies 9, 4, 1, .25 ; total 14.25
    assert vo.ref_find_n_decorrelation(s, 0.79) == 1          # 0.79^2 * 14.25 = 8.89 <= 9
    assert vo.ref_find_n_decorrelation(s, 0.80) == 2          # 0.64 * 14.25 = 9.12 > 9
    assert vo.ref_find_n_decorrelation(s, 1.0) == 4
    assert vo.ref_find_n_decorrelation(s, 1e-6) == 1


def test_rank_precedence_and_falsy_options(golden):
    a = golden["syn_64x64/A"]
    assert len(vo.ref_apply_svd(a, decorrelation=0.9, compressionrank=3)[1]) == 3      # fixed rank wins
    assert len(vo.ref_apply_svd(a, decorrelation=0, compressionrank=0)[1]) == 64        # both falsy -> full
    assert len(vo.ref_apply_svd(a)[1]) == 64


def test_reconstruction_errors_match_reference(golden, golden_cases):
    for c in golden_cases:
        a = golden[f"{c}/A"]
        for k, e_ref in zip(golden[f"{c}/ks"], golden[f"{c}/recon_err"]):
            rec, s, kk = vo.roundtrip(a, compressionrank=int(k))
            assert kk == k and rec.dtype == np.complex64
            e = np.linalg.norm(a.astype(np.complex128) - rec.astype(np.complex128))
            assert abs(e - e_ref) <= 1e-6 * e_ref + 1e-6 * np.linalg.norm(a), (c, k)


def test_stored_reconstruction_bitwise(golden, golden_cases):
    for c in golden_cases:
        a = golden[f"{c}/A"]
        k = min(2, min(a.shape))
        u, s, vt = vo.ref_apply_svd(a, decorrelation=0.9, compressionrank=k)
        np.testing.assert_array_equal(vo.ref_reconstruct_vis(u, s, vt), golden[f"{c}/recon_k2"])
        np.testing.assert_array_equal(vo.ref_reconstruct_vis(u, s.reshape(-1, 1), vt), golden[f"{c}/recon_k2"])


def test_unstack(golden):
    a = golden["ms_bl12_diag/A"]
    u, s, vt = vo.ref_apply_svd(a, compressionrank=4)
    parts = vo.ref_unstack_vis(vo.ref_reconstruct_vis(u, s, vt), 360)
    assert len(parts) == 2
    np.testing.assert_array_equal(parts[0], golden["ms_bl12_diag/unstack0"])
    np.testing.assert_array_equal(parts[1], golden["ms_bl12_diag/unstack1"])


def test_svd_flip_convention():
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((6, 9)) + 1j * rng.standard_normal((6, 9))).astype(np.complex64)
    u, s, vt = vo.ref_svd(a)
    sums = vt.sum(axis=1)
    assert np.all((sums.real > 0) | ((sums.real == 0) & (sums.imag >= 0)))
    np.testing.assert_allclose((u * s) @ vt, a, atol=1e-5)


@pytest.mark.parametrize("shape", [(1, 40), (40, 1)])
def test_degenerate_shapes(golden, shape):
    a = golden[f"syn_{shape[0]}x{shape[1]}/A"]
    rec, s, k = vo.roundtrip(a, decorrelation=0.99)
    assert k == 1
    np.testing.assert_allclose(rec, a, atol=1e-5)
