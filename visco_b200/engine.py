"""Host-side engine: one ``Engine`` = one ``vk_handle`` = one GPU. torch is used only for device memory and
streams; every arithmetic step is a call into libvisco_b200.so (include/visco_b200.h).

The batched methods are what the rewritten L2 loops of the reference call with a whole dask batch of matrices
(reference visco/compress_ms.py:571-697 and visco/decompress_ms.py:196-213) instead of one task per matrix.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("visco_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = int(device)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        torch.zeros(1, device=f"cuda:{self.device}")  # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.vk_create(C.byref(h), self.device)
        if rc != _lib.VK_OK:
            raise RuntimeError(f"vk_create failed with status {rc} (is this an sm_100 GPU?)")
        self.h = h
        self._lock = threading.Lock()
        # stream of the host-buffer entry points: None = the legacy default stream; a torch.cuda.Stream lets several
        # engines (one per host thread) overlap their copies and kernels on one GPU
        self.host_stream = None

    # ------------------------------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "h", None):
            self.lib.vk_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        _lib.raise_for_status(self.lib, self.h, rc, what)

    def set_option(self, key: str, value: float):
        self._check(self.lib.vk_set_option(self.h, key.encode(), float(value)), f"vk_set_option({key})")

    def _bind_stream(self):
        torch = _torch()
        self.lib.vk_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def sync(self):
        self._check(self.lib.vk_sync(self.h), "vk_sync")

    @property
    def launch_count(self) -> int:
        return int(self.lib.vk_launch_count(self.h))

    def last_stage_ms(self):
        t = (C.c_float * 6)()
        self.lib.vk_last_stage_ms(self.h, t)
        return dict(zip(("gram", "jacobi", "select", "factors", "small", "total"), [float(x) for x in t]))

    def last_eig_ms(self):
        t = (C.c_float * 5)()
        self.lib.vk_last_eig_ms(self.h, t)
        return dict(zip(("tridiag", "leading_pairs", "ql", "reflectors", "rotations"), [float(x) for x in t]))

    def uses_small_path(self, m, n) -> bool:
        return bool(self.lib.vk_uses_small_path(int(m), int(n)))

    def gram_uses_tcgen05(self, m, n) -> bool:
        return bool(self.lib.vk_gram_uses_tcgen05(int(m), int(n), 0 if m <= n else 1))

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    @staticmethod
    def rank_bound(m, n, compressionrank=None, decorrelation=None) -> int:
        r = min(m, n)
        return min(int(compressionrank), r) if compressionrank else r

    # ------------------------------------------------------------------------------------------ device API
    def compress(self, A, decorrelation=None, compressionrank=None, kmax=None, out=None):
        """A: torch complex64 CUDA tensor [B, m, n] (contiguous). Returns (U, S, Vt, ranks, stats) torch tensors
        laid out as include/visco_b200.h describes. Rank rule as reference apply_svd (compress_ms.py:352-357)."""
        torch = _torch()
        assert A.is_cuda and A.dtype == torch.complex64 and A.dim() == 3 and A.is_contiguous()
        B, m, n = A.shape
        if kmax is None:
            kmax = self.rank_bound(m, n, compressionrank, decorrelation)
        dev = A.device
        if out is None:
            U = torch.empty((B, m, kmax), dtype=torch.complex64, device=dev)
            S = torch.empty((B, kmax), dtype=torch.float32, device=dev)
            Vt = torch.empty((B, kmax, n), dtype=torch.complex64, device=dev)
            ranks = torch.empty((B,), dtype=torch.int32, device=dev)
            stats = torch.empty((B, 4), dtype=torch.float32, device=dev)
        else:
            U, S, Vt, ranks, stats = out
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_compress_batched(self.h, self._ptr(A), B, m, n, int(compressionrank or 0),
                                              float(decorrelation or 0.0), int(kmax), self._ptr(U), self._ptr(S),
                                              self._ptr(Vt), self._ptr(ranks), self._ptr(stats), C.c_void_p(0), 0)
        self._check(rc, "vk_compress_batched")
        return U, S, Vt, ranks, stats

    def reconstruct(self, U, S, Vt, ranks=None, out=None):
        """out[b] = (U[b] * S[b]) @ Vt[b] on device (reference reconstruct_vis, decompress_ms.py:107-131)."""
        torch = _torch()
        B, m, kmax = U.shape
        n = Vt.shape[2]
        assert U.is_contiguous() and Vt.is_contiguous() and S.is_contiguous()
        if out is None:
            out = torch.empty((B, m, n), dtype=torch.complex64, device=U.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_reconstruct_batched(self.h, self._ptr(U), self._ptr(S), self._ptr(Vt), self._ptr(ranks),
                                                 B, m, n, kmax, self._ptr(out))
        self._check(rc, "vk_reconstruct_batched")
        return out

    def find_n_decorrelation(self, S, decorrelation: float):
        """S: torch float32 CUDA [B, r] descending -> int32 ranks [B] (reference compress_ms.py:295-319)."""
        torch = _torch()
        B, r = S.shape
        ranks = torch.empty((B,), dtype=torch.int32, device=S.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_find_n_decorrelation_batched(self.h, self._ptr(S), B, r, float(decorrelation),
                                                          self._ptr(ranks))
        self._check(rc, "vk_find_n_decorrelation_batched")
        return ranks

    def synth_fill(self, A, nbl_local, ncorr, bl_offset=0, nbl_total=None, seed=20261018):
        """Fill A [nbl_local*ncorr, m, n] with the SURVEY section 8d synthetic visibilities (device generator)."""
        B, m, n = A.shape
        assert B == nbl_local * ncorr
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_synth_fill(self.h, self._ptr(A), nbl_local, ncorr, m, n, bl_offset,
                                        nbl_total or nbl_local, seed)
        self._check(rc, "vk_synth_fill")
        return A

    # layout kernels on either side of the path
    def gather_baselines(self, data, row_idx, corr_sel, stack=1, out=None):
        """data: CUDA complex64 [row, chan, corr]; row_idx: CUDA int32 [nbl, m] (-1 = padding); corr_sel: CUDA int32
        [nbl, ncs] correlation planes per entry. Returns the cube [nbl * ncs / stack, stack * m, chan] (reference compress_ms.py:591-664)."""
        torch = _torch()
        nrow, nchan, ncorr = data.shape
        nbl, m = row_idx.shape
        ncs = corr_sel.shape[1]
        assert data.is_contiguous() and row_idx.dtype == torch.int32 and corr_sel.dtype == torch.int32
        assert corr_sel.shape[0] == nbl and row_idx.is_contiguous() and corr_sel.is_contiguous()
        if out is None:
            out = torch.zeros((nbl * ncs // stack, stack * m, nchan), dtype=torch.complex64, device=data.device)
        self._check_layout(row_idx, corr_sel, nrow, ncorr)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_gather_baselines(self.h, self._ptr(data), nchan, ncorr, self._ptr(row_idx), nbl, m,
                                              self._ptr(corr_sel), ncs, stack, self._ptr(out))
        self._check(rc, "vk_gather_baselines")
        return out

    def _check_layout(self, row_idx, corr_sel, nrow, ncorr):
        """ValueError when a row index is outside [-1, nrow) or a correlation plane outside [0, ncorr): the reference's
        numpy indexing raises IndexError there (decompress_ms.py:216-232); the kernels would touch foreign memory."""
        bad = C.c_int32(0)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_check_layout_indices(self.h, self._ptr(row_idx), row_idx.numel(), int(nrow),
                                                  self._ptr(corr_sel), corr_sel.numel(), int(ncorr), C.byref(bad))
        self._check(rc, "vk_check_layout_indices")

    def scatter_baselines(self, cube, data, row_idx, corr_sel, stack=1):
        """Inverse of gather_baselines: writes the matrices of `cube` into data[row, chan, corr] in place
        (reference decompress_ms.py:216-232)."""
        torch = _torch()
        nrow, nchan, ncorr = data.shape
        nbl, m = row_idx.shape
        ncs = corr_sel.shape[1]
        assert corr_sel.shape[0] == nbl and row_idx.is_contiguous() and corr_sel.is_contiguous()
        assert cube.is_contiguous() and data.is_contiguous() and cube.shape == (nbl * ncs // stack, stack * m, nchan)
        self._check_layout(row_idx, corr_sel, nrow, ncorr)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_scatter_baselines(self.h, self._ptr(cube), nchan, ncorr, self._ptr(row_idx), nbl, m,
                                               self._ptr(corr_sel), ncs, stack, self._ptr(data))
        self._check(rc, "vk_scatter_baselines")
        return data

    # flags
    def packbits(self, flags):
        """np.packbits(flags, axis=None) on device: flags is a CUDA bool/uint8 tensor of any shape."""
        torch = _torch()
        f = flags.contiguous().view(torch.uint8).reshape(-1)
        out = torch.empty(((f.numel() + 7) // 8,), dtype=torch.uint8, device=f.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_packbits(self.h, self._ptr(f), f.numel(), self._ptr(out))
        self._check(rc, "vk_packbits")
        return out

    def unpackbits(self, packed, count):
        """np.unpackbits(packed, count=count) on device -> CUDA uint8 tensor [count]."""
        torch = _torch()
        out = torch.empty((int(count),), dtype=torch.uint8, device=packed.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_unpackbits(self.h, self._ptr(packed.contiguous()), int(count), self._ptr(out))
        self._check(rc, "vk_unpackbits")
        return out

    def flag_replace(self, data, flags, model=None, value=0j):
        """In place: data = where(flags, model or value, data) (reference compress_ms.py:530-562)."""
        torch = _torch()
        f = flags.contiguous().view(torch.uint8)
        assert data.is_contiguous() and f.numel() == data.numel() and (model is None or model.shape == data.shape)
        v = complex(value)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_flag_replace(self.h, self._ptr(data), self._ptr(f), self._ptr(model.contiguous()) if model is not None else None,
                                          float(v.real), float(v.imag), data.numel())
        self._check(rc, "vk_flag_replace")
        return data

    # stage-level (tests / profiling)
    def gram(self, A, impl=0):
        torch = _torch()
        B, m, n = A.shape
        side = 0 if m <= n else 1
        r = min(m, n)
        W = torch.empty((B, r, r), dtype=torch.complex64, device=A.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_gram_batched(self.h, self._ptr(A), B, m, n, side, impl, self._ptr(W))
        self._check(rc, "vk_gram_batched")
        return W

    def eigh_jacobi(self, W):
        torch = _torch()
        B, r, _ = W.shape
        lam = torch.empty((B, r), dtype=torch.float32, device=W.device)
        info = torch.empty((B, 2), dtype=torch.int32, device=W.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_eigh_jacobi_batched(self.h, self._ptr(W), B, r, self._ptr(lam), self._ptr(info))
        self._check(rc, "vk_eigh_jacobi_batched")
        return lam, info

    def svd_small(self, A):
        torch = _torch()
        B, m, n = A.shape
        r = min(m, n)
        U = torch.empty((B, m, r), dtype=torch.complex64, device=A.device)
        S = torch.empty((B, r), dtype=torch.float32, device=A.device)
        Vt = torch.empty((B, r, n), dtype=torch.complex64, device=A.device)
        info = torch.empty((B, 2), dtype=torch.int32, device=A.device)
        with self._lock:
            self._bind_stream()
            rc = self.lib.vk_svd_jacobi_small_batched(self.h, self._ptr(A), B, m, n, self._ptr(U), self._ptr(S),
                                                      self._ptr(Vt), self._ptr(info))
        self._check(rc, "vk_svd_jacobi_small_batched")
        return U, S, Vt, info

    # ------------------------------------------------------------------------------------------ host API
    def compress_host(self, A: np.ndarray, decorrelation=None, compressionrank=None, out=None):
        """numpy [B, m, n] complex64 in, padded numpy factors + ranks + stats out (vk_compress_host).
        `out` = (U, S, Vt, ranks, stats) lets the caller supply (pinned) result buffers."""
        A = np.ascontiguousarray(A, dtype=np.complex64)
        B, m, n = A.shape
        kmax = self.rank_bound(m, n, compressionrank, decorrelation)
        if out is None:
            U = np.empty((B, m, kmax), np.complex64)
            S = np.empty((B, kmax), np.float32)
            Vt = np.empty((B, kmax, n), np.complex64)
            ranks = np.empty((B,), np.int32)
            stats = np.empty((B, 4), np.float32)
        else:
            U, S, Vt, ranks, stats = out
            assert U.shape == (B, m, kmax) and S.shape == (B, kmax) and Vt.shape == (B, kmax, n)
            assert all(x.flags.c_contiguous for x in out)
        with self._lock:
            self.lib.vk_set_stream(self.h, C.c_void_p(self.host_stream.cuda_stream if self.host_stream is not None else 0))
            rc = self.lib.vk_compress_host(self.h, A.ctypes.data, B, m, n, int(compressionrank or 0),
                                           float(decorrelation or 0.0), kmax, U.ctypes.data, S.ctypes.data,
                                           Vt.ctypes.data, ranks.ctypes.data, stats.ctypes.data)
        self._check(rc, "vk_compress_host")
        return U, S, Vt, ranks, stats

    def reconstruct_host(self, U: np.ndarray, S: np.ndarray, Vt: np.ndarray, ranks=None, out=None):
        U = np.ascontiguousarray(U, dtype=np.complex64)
        S = np.ascontiguousarray(S, dtype=np.float32)
        Vt = np.ascontiguousarray(Vt, dtype=np.complex64)
        B, m, kmax = U.shape
        n = Vt.shape[2]
        if S.shape != (B, kmax) or Vt.shape[:2] != (B, kmax):
            raise ValueError(f"inconsistent factor shapes U{U.shape} S{S.shape} Vt{Vt.shape}")
        if out is None:
            out = np.empty((B, m, n), np.complex64)
        assert out.shape == (B, m, n) and out.dtype == np.complex64 and out.flags.c_contiguous
        rp = None
        if ranks is not None:
            rp = np.ascontiguousarray(ranks, dtype=np.int32)
        with self._lock:
            self.lib.vk_set_stream(self.h, C.c_void_p(self.host_stream.cuda_stream if self.host_stream is not None else 0))
            rc = self.lib.vk_reconstruct_host(self.h, U.ctypes.data, S.ctypes.data, Vt.ctypes.data,
                                              rp.ctypes.data if rp is not None else None, B, m, n, kmax,
                                              out.ctypes.data)
        self._check(rc, "vk_reconstruct_host")
        return out


_engines = {}
_engines_lock = threading.Lock()


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine for a device (default: torch's current device)."""
    if device is None:
        torch = _torch()
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    with _engines_lock:
        if device not in _engines:
            _engines[device] = Engine(device)
        return _engines[device]
