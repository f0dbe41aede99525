"""CPU: the C-ABI library loads and exports every symbol include/visco_b200.h declares; the ctypes table matches the
header; host-side argument handling that needs no GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "visco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _header_functions()
    for need in ("vk_compress_batched", "vk_reconstruct_batched", "vk_find_n_decorrelation_batched",
                 "vk_compress_host", "vk_reconstruct_host", "vk_gram_batched", "vk_eigh_jacobi_batched",
                 "vk_svd_jacobi_small_batched", "vk_synth_fill", "vk_create", "vk_destroy", "vk_last_error"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from visco_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    for name in _header_functions():
        assert hasattr(lib, name), f"{name} declared in include/visco_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == _header_functions()
    assert b"sm_100a" in lib.vk_version()


def test_path_selection_is_pure_host_logic():
    from visco_b200 import _lib
    lib = _lib.load()
    assert lib.vk_uses_small_path(64, 64) == 0          # BASELINE config 4: Gram path by default ("small_impl" = 1: Jacobi)
    assert lib.vk_uses_small_path(32, 64) == 1
    assert lib.vk_uses_small_path(360, 16) == 1         # sample MS
    assert lib.vk_uses_small_path(16, 360) == 1
    assert lib.vk_uses_small_path(256, 1024) == 0       # config 2 -> Gram path
    assert lib.vk_uses_small_path(64, 4096) == 0        # r <= 64 but too long for shared memory -> Gram path
    assert lib.vk_uses_small_path(65, 65) == 0
    assert lib.vk_workspace_bytes(None, 112, 256, 1024, 8) >= 112 * 256 * 256 * 8


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from visco_b200.compress_ms import apply_svd
    with pytest.raises(RuntimeError):
        apply_svd(np.zeros((4, 4), np.complex64), compressionrank=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "visco_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_every_library_option_is_documented_in_the_header():
    """vk_set_option's table (csrc/api.cu) and the option list in include/visco_b200.h must not drift apart."""
    import re
    api = open(os.path.join(ROOT, "visco_b200", "csrc", "api.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "visco_b200.h")).read()
    opts = re.findall(r'k == "([a-z0-9_]+)"', api)
    assert len(opts) >= 30 and len(set(opts)) == len(opts)
    missing = [o for o in opts if f'"{o}"' not in hdr]
    assert not missing, f"options not described in include/visco_b200.h: {missing}"
