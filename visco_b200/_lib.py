"""ctypes binding of libvisco_b200.so (include/visco_b200.h). There is NO fallback: if the CUDA library is
missing, importing this module works (so host logic can be tested) but every compute call raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvisco_b200.so")

VK_OK, VK_EINVAL, VK_ENOMEM, VK_ECUDA, VK_ENOCONV, VK_ENONFINITE = range(6)

_vp, _i, _f, _d, _sz, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint64

# name -> (restype, argtypes): exactly the declarations of include/visco_b200.h
SIGNATURES = {
    "vk_create": (_i, [C.POINTER(_vp), _i]),
    "vk_destroy": (_i, [_vp]),
    "vk_last_error": (C.c_char_p, [_vp]),
    "vk_version": (C.c_char_p, []),
    "vk_set_stream": (_i, [_vp, _vp]),
    "vk_sync": (_i, [_vp]),
    "vk_set_option": (_i, [_vp, C.c_char_p, C.c_double]),
    "vk_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "vk_compress_batched": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "vk_reconstruct_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "vk_compress_host": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "vk_reconstruct_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "vk_find_n_decorrelation_batched": (_i, [_vp, _vp, _i, _i, _d, _vp]),
    "vk_gram_batched": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "vk_eigh_jacobi_batched": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "vk_svd_jacobi_small_batched": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "vk_uses_small_path": (_i, [_i, _i]),
    "vk_gram_uses_tcgen05": (_i, [_i, _i, _i]),
    "vk_gather_baselines": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "vk_scatter_baselines": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "vk_check_layout_indices": (_i, [_vp, _vp, _sz, _i, _vp, _sz, _i, C.POINTER(C.c_int32)]),
    "vk_packbits": (_i, [_vp, _vp, _sz, _vp]),
    "vk_unpackbits": (_i, [_vp, _vp, _sz, _vp]),
    "vk_flag_replace": (_i, [_vp, _vp, _vp, _vp, _f, _f, _sz]),
    "vk_synth_fill": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _u64]),
    "vk_launch_count": (C.c_int64, [_vp]),
    "vk_last_stage_ms": (_i, [_vp, C.POINTER(_f)]),
    "vk_last_eig_ms": (_i, [_vp, C.POINTER(_f)]),
}

_lib = None


class ViscoLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed to load it). Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ViscoLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m visco_b200._build` (or __graft_entry__.build()). "
            "visco_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def raise_for_status(lib, handle, rc, what):
    """Error convention of the boundary: status code -> the exception the reference raises."""
    if rc == VK_OK:
        return
    msg = lib.vk_last_error(handle).decode() if handle else "no handle"
    text = f"{what}: {msg}"
    if rc in (VK_EINVAL, VK_ENONFINITE):
        raise ValueError(text)
    if rc == VK_ENOMEM:
        raise MemoryError(text)
    if rc == VK_ENOCONV:
        import numpy as np
        raise np.linalg.LinAlgError(text)
    raise RuntimeError(text)
