"""Generate tests/golden/*.npz by running the REFERENCE's own function bodies in this container.

Run once, here (needs /root/reference; never runs on the GPU box):   python tests/golden/make_golden.py

The reference package cannot be imported (dask, xarray, zarr, dask-ms, omegaconf are absent), so this script
  1. parses visco/compress_ms.py and visco/decompress_ms.py with `ast` and exec()s ONLY the four hot-path
     functions (find_n_decorrelation, apply_svd, unstack_vis, reconstruct_vis) — their bodies unchanged —
  2. in a namespace whose `da` is a small numpy-backed stand-in for `dask.array` providing exactly the calls
     those bodies make: da.Array, da.from_array, da.linalg.svd (single-chunk branch of dask 2024.10.0:
     np.linalg.svd(full_matrices=False) + svd_flip), da.sum, da.cumsum, and `.compute()`.
Inputs: matrices decoded from the reference's sample Measurement Set (tests/data/sim-visco-kat7.ms, DATA
column, TiledShapeStMan file table.f2_TSM1) plus seeded synthetic matrices. Inputs and outputs are stored.
"""
import ast
import os
import sys

import numpy as np

REF = os.environ.get("VISCO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------ numpy stand-in for dask.array
class _Arr(np.ndarray):
    def compute(self):
        a = np.asarray(self)
        return a[()] if a.ndim == 0 else a


def _wrap(x):
    return np.asarray(x).view(_Arr)


def _svd_flip(u, v):
    dtype = v.dtype
    signs = np.sum(v, axis=1, keepdims=True).T
    signs = 2.0 * ((signs >= 0) - 0.5).astype(dtype)
    return u * signs, v * signs.T


class _Linalg:
    @staticmethod
    def svd(a):
        u, s, v = np.linalg.svd(np.asarray(a), full_matrices=False)
        u, v = _svd_flip(u, v)
        return _wrap(u), _wrap(s), _wrap(v)


class _Da:
    Array = _Arr
    linalg = _Linalg

    @staticmethod
    def from_array(x, chunks=None):
        return _wrap(x)

    @staticmethod
    def sum(x):
        return _wrap(np.sum(np.asarray(x)))

    @staticmethod
    def cumsum(x):
        return _wrap(np.cumsum(np.asarray(x)))


def load_reference_functions():
    ns = {"np": np, "da": _Da}
    wanted = {"visco/compress_ms.py": ["find_n_decorrelation", "apply_svd"],
              "visco/decompress_ms.py": ["unstack_vis", "reconstruct_vis"]}
    for rel, names in wanted.items():
        src = open(os.path.join(REF, rel)).read()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                code = compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, rel), "exec")
                exec(code, ns)
    return ns


# ------------------------------------------------------------------ sample MS decode (no casacore)
def read_sample_ms_data():
    """DATA column of tests/data/sim-visco-kat7.ms: 76 tiles of [100 rows][16 chan][4 corr] complex64 LE;
    7560 valid rows = 360 timeslots x 21 baselines (time-major)."""
    path = os.path.join(REF, "tests/data/sim-visco-kat7.ms/table.f2_TSM1")
    raw = np.fromfile(path, dtype="<c8")
    cube = raw.reshape(76 * 100, 16, 4)[:7560]
    return cube.reshape(360, 21, 16, 4)  # [time][baseline][chan][corr]


def synth(m, n, seed, gain=1.0, nsrc=6, rate=8.0):
    rng = np.random.default_rng(seed)
    t = np.arange(m)[:, None] / m
    nu = np.arange(n)[None, :] / max(n, 1)
    a = np.zeros((m, n), np.complex128)
    for _ in range(nsrc):
        rho = rng.uniform(-rate, rate)
        phi = rng.uniform(0, 2 * np.pi)
        a += np.exp(1j * (2 * np.pi * rho * t * (1 + 0.2 * nu) + phi))
    a *= gain
    a += (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
    return a.astype(np.complex64)


def main():
    fn = load_reference_functions()
    apply_svd, find_n, recon, unstack = (fn["apply_svd"], fn["find_n_decorrelation"],
                                         fn["reconstruct_vis"], fn["unstack_vis"])
    ms = read_sample_ms_data()
    cases = {
        "ms_bl00_XX": ms[:, 0, :, 0].copy(),        # 360 x 16 (m > n)
        "ms_bl07_YY": ms[:, 7, :, 3].copy(),
        "ms_bl20_XY": ms[:, 20, :, 1].copy(),
        "ms_bl12_diag": np.vstack([ms[:, 12, :, 0], ms[:, 12, :, 3]]),   # corr-optimized stack 720 x 16
        "syn_64x64": synth(64, 64, 1),
        "syn_48x200": synth(48, 200, 2),
        "syn_200x48": synth(200, 48, 3),
        "syn_128x256": synth(128, 256, 4),
        "syn_96x160_weak": synth(96, 160, 5, gain=0.01),
        "syn_33x70": synth(33, 70, 6),
        "syn_1x40": synth(1, 40, 7),
        "syn_40x1": synth(40, 1, 8),
        "syn_rank3_32x64": (synth(32, 3, 9, nsrc=0) @ synth(3, 64, 10, nsrc=0)).astype(np.complex64),
    }
    decs = [0.5, 0.9, 0.95, 0.99, 0.999, 1.0]
    out = {}
    for name, a in cases.items():
        a = np.ascontiguousarray(a, dtype=np.complex64)
        u, s, vt = apply_svd(a)                       # neither option -> full rank
        u, s, vt = np.asarray(u), np.asarray(s), np.asarray(vt)
        assert s.dtype == np.float32 and u.dtype == np.complex64
        out[f"{name}/A"] = a
        out[f"{name}/S"] = s
        out[f"{name}/n_dec"] = np.array([find_n(_wrap(s), d) for d in decs], np.int64)
        r = len(s)
        ks = sorted({1, min(2, r), min(8, r), max(1, r // 2), r})
        errs = []
        for k in ks:
            uk, sk, vk = apply_svd(a, compressionrank=k)
            rec = np.asarray(recon(uk, sk, vk))
            assert rec.dtype == np.complex64
            errs.append(np.linalg.norm(a.astype(np.complex128) - rec.astype(np.complex128)))
        out[f"{name}/ks"] = np.array(ks, np.int64)
        out[f"{name}/recon_err"] = np.array(errs, np.float64)
        # precedence: fixed rank wins over decorrelation (compress_ms.py:352-355)
        uk, sk, vk = apply_svd(a, decorrelation=0.9, compressionrank=min(2, r))
        assert len(np.asarray(sk)) == min(2, r)
        # one stored reconstruction (k = min(2, r)) + S column-vector form accepted (decompress_ms.py:125-126)
        rec = np.asarray(recon(uk, sk, vk))
        rec2 = np.asarray(recon(uk, _wrap(np.asarray(sk).reshape(-1, 1)), vk))
        assert np.array_equal(rec, rec2)
        out[f"{name}/recon_k2"] = rec
    # unstack_vis on the stacked case
    a = cases["ms_bl12_diag"]
    uk, sk, vk = apply_svd(a, compressionrank=4)
    parts = unstack(np.asarray(recon(uk, sk, vk)), 360)
    out["ms_bl12_diag/unstack0"] = np.asarray(parts[0])
    out["ms_bl12_diag/unstack1"] = np.asarray(parts[1])
    out["decs"] = np.array(decs)
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), os.path.getsize(os.path.join(HERE, "hotpath_golden.npz")), "bytes")
    print("numpy", np.__version__)


if __name__ == "__main__":
    sys.exit(main())
