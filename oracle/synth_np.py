"""TEST / BENCH INFRASTRUCTURE ONLY: numpy twin of the synthetic visibility model, used by
`bench.py --impl reference` so the CPU arm needs nothing from the CUDA library. Same model and parameters as
visco_b200/csrc/stages.cu:synth_kernel; the random streams differ (numpy PCG64 here), the statistics do not."""
import numpy as np


def synth_cube(nbl, ncorr, m, n, bl_offset=0, nbl_total=None, seed=20261018, nsrc=10):
    nbl_total = nbl_total or nbl
    out = np.empty((nbl * ncorr, m, n), np.complex64)
    t = (np.arange(m, dtype=np.float64) / m)[:, None]
    fv = 1.0 + 0.2 * (np.arange(n, dtype=np.float64) / n)[None, :]
    for bl in range(nbl):
        gbl = bl_offset + bl
        rng = np.random.default_rng([seed, gbl])
        R = 30.0 * (gbl + 1) / nbl_total
        rho = rng.uniform(-R, R, nsrc)
        phi = rng.uniform(0, 2 * np.pi, nsrc)
        sky = np.zeros((m, n), np.complex128)
        for s in range(nsrc):
            sky += np.exp(1j * (2 * np.pi * rho[s] * t * fv + phi[s]))
        for c in range(ncorr):
            gain = 1.0 if c in (0, ncorr - 1) else 0.01
            noise = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
            out[bl * ncorr + c] = (gain * sky + noise).astype(np.complex64)
    return out
