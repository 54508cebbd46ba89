"""ctypes binding of libcora_b200.so (the C ABI declared in include/cora_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
PyTorch is used only to own device buffers and streams (``tensor.data_ptr()``).
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcora_b200.so")

ALM_PACKED = 0
ALM_PANEL = 1

_c = ctypes
_vp, _i, _ll, _d, _ull = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_double, _c.c_ulonglong

# name -> (restype, argtypes); mirrors include/cora_b200.h one to one
SIGNATURES = {
    "cora_b200_version": (_i, []),
    "cora_b200_last_error": (_c.c_char_p, []),
    "cora_b200_launch_count": (_ll, []),
    "cora_b200_fp64_peak": (_i, [_d, _c.POINTER(_d), _vp]),
    "cora_b200_timing_enable": (_i, [_i]),
    "cora_b200_timing_kinds": (_i, []),
    "cora_b200_timing_name": (_c.c_char_p, [_i]),
    "cora_b200_timing_read": (_i, [_c.POINTER(_d), _c.POINTER(_ll), _i]),
    "cora_b200_timing_trace": (_i, [_c.POINTER(_i), _c.POINTER(_d), _c.POINTER(_d), _i]),
    "cora_b200_sht_plan_create": (_i, [_i, _i, _c.POINTER(_vp)]),
    "cora_b200_sht_plan_destroy": (_i, [_vp]),
    "cora_b200_alm2map_workspace_bytes": (_ll, [_vp, _i, _i]),
    "cora_b200_alm2map": (_i, [_vp, _vp, _i, _ll, _i, _vp, _vp, _ll, _vp]),
    "cora_b200_alm2map_spin2": (_i, [_vp, _vp, _vp, _i, _ll, _i, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_alm_panel_to_dense": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp]),
    "cora_b200_alm_dense_to_panel": (_i, [_vp, _i, _i, _vp, _ll, _i, _vp]),
    "cora_b200_cl_fill_sck": (_i, [_d, _d, _d, _d, _d, _d, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "cora_b200_ps_table_21cm": (_i, [_vp, _vp, _vp, _i, _d, _vp, _vp, _ll, _vp]),
    "cora_b200_ps_table_21cm_bytes": (_ll, []),
    "cora_b200_ps_table_21cm_workspace_bytes": (_ll, []),
    "cora_b200_cl_fill_21cm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "cora_b200_cl_symmetrize": (_i, [_vp, _i, _i, _vp]),
    "cora_b200_cl_fill_21cm_ntiles": (_ll, [_i]),
    "cora_b200_cl_romberg_reduce": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "cora_b200_aps_sck_points": (_i, [_d, _d, _d, _d, _d, _d, _vp, _vp, _vp, _ll, _vp, _vp]),
    "cora_b200_aps_21cm_points": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp]),
    "cora_b200_ps_table_21cm_gather": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "cora_b200_root_workspace_bytes": (_ll, [_i, _i]),
    "cora_b200_root_batched": (_i, [_vp, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_eigh_workspace_bytes": (_ll, [_i, _i]),
    "cora_b200_eigh_batched": (_i, [_vp, _i, _i, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_draw_apply_workspace_bytes": (_ll, [_i, _i, _i]),
    "cora_b200_draw_bytes": (_ll, [_vp, _i, _i]),
    "cora_b200_draw": (_i, [_vp, _i, _i, _ull, _i, _vp, _ll, _vp]),
    "cora_b200_draw_apply": (_i, [_vp, _vp, _vp, _i, _i, _i, _ull, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _ll, _vp]),
    "cora_b200_draw_apply_slabs": (_i, [_vp, _vp, _vp, _i, _i, _i, _ull, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_alm_slabs_to_panel": (_i, [_vp, _vp, _i, _i, _vp, _ll, _i, _vp]),
    "cora_b200_map2alm_workspace_bytes": (_ll, [_vp, _i]),
    "cora_b200_map2alm": (_i, [_vp, _vp, _i, _vp, _i, _vp, _ll, _i, _vp, _ll, _vp]),
    "cora_b200_map2alm_spin2_workspace_bytes": (_ll, [_vp, _i]),
    "cora_b200_map2alm_spin2": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _ll, _i, _vp, _ll, _vp]),
    "cora_b200_map_sub": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "cora_b200_peer_alloc": (_i, [_ll, _c.POINTER(_vp), _c.c_char_p]),
    "cora_b200_peer_free": (_i, [_vp]),
    "cora_b200_peer_open": (_i, [_c.c_char_p, _c.POINTER(_vp)]),
    "cora_b200_peer_close": (_i, [_vp]),
    "cora_b200_peer_barrier": (_i, [_vp, _i, _i, _ull, _d, _vp, _i, _vp]),
    "cora_b200_cl_fill_21cm_tiles": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _i, _i, _vp, _vp, _vp, _vp]),
    "cora_b200_draw_apply_peers": (_i, [_vp, _vp, _vp, _i, _i, _i, _ull, _i, _vp, _ll, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_alm2map_strided": (_i, [_vp, _vp, _i, _ll, _i, _vp, _ll, _vp, _ll, _vp]),
    "cora_b200_alm2map_spin2_strided": (_i, [_vp, _vp, _vp, _i, _ll, _i, _vp, _vp, _ll, _vp, _ll, _vp]),
    "cora_b200_diag_max": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "cora_b200_root_multi_workspace_bytes": (_ll, [_i, _i, _i]),
    "cora_b200_root_batched_multi": (_i, [_vp, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "cora_b200_alm_scale_l": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp]),
    "cora_b200_set_legendre_ws": (_i, [_i]),
    "cora_b200_corr_bins": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "cora_b200_legendre_table": (_i, [_vp, _vp, _i, _i, _vp, _ll, _vp]),
    "cora_b200_dgemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _ll, _ll, _ll, _i, _vp]),
}

_lib = None


class CoraB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoraB200Error(
            "libcora_b200.so not found at %s -- build it with `python -m cora_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cora_b200_last_error()
        raise CoraB200Error("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def call(name, *args):
    """Call an int-returning ABI function and raise on a non-zero code."""
    check(getattr(load(), name)(*args), name)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise CoraB200Error("cora_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    return torch


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array as c_void_p."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def ptr_off(t, nbytes):
    """Device pointer of a torch tensor advanced by ``nbytes``."""
    return ctypes.c_void_p(t.data_ptr() + int(nbytes))


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)
