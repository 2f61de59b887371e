"""ctypes binding of ``libpydmrg_b200.so`` (the C ABI in ``include/pydmrg_b200.h``).

This is the reference-side stub a PyDMRG maintainer would add: plain
pointers and sizes, no torch types in any signature.  PyTorch is used above this layer only
to own device buffers and streams.  There is NO CPU fallback: if the shared library is
missing or fails to load, importing the product modules raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_ubyte,
                    c_void_p)

F64 = 0
C128 = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpydmrg_b200.so")


class PdmrgError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pydmrg_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    vp, ll, i, d, sz = c_void_p, c_longlong, c_int, c_double, c_size_t
    sig = {
        "pdmrg_last_error": (c_char_p, []),
        "pdmrg_version": (i, []),
        "pdmrg_launch_count": (ll, []),
        "pdmrg_last_gemm_path": (i, []),
        "pdmrg_set_force_generic_gemm": (None, [i]),
        "pdmrg_gemm_nt": (i, [i, i, i, i, i, vp, ll, ll, vp, ll, ll, i, vp, ll, ll, d, d, i, vp]),
        "pdmrg_gemm_nt_batched2": (i, [i, i, i, i, i, i, POINTER(i), vp, ll, ll, ll, vp, ll, ll, ll, i, vp, ll, ll, ll, d, d, i, vp]),
        "pdmrg_set_force_tile_rows": (None, [i]),
        "pdmrg_set_streamk_mode": (None, [i]),
        "pdmrg_last_gemm_streamk": (i, []),
        "pdmrg_permute4": (i, [i, vp, vp, i, i, i, i, ll, ll, ll, ll, i, vp, i, vp]),
        "pdmrg_heff_plan_create": (i, [POINTER(vp), i, i, i, i, i, POINTER(i), POINTER(i), POINTER(vp)]),
        "pdmrg_heff_plan_create_sharded": (i, [POINTER(vp), i, i, i, i, i, POINTER(i), POINTER(i), POINTER(vp), POINTER(vp)]),
        "pdmrg_heff_plan_create_ketsplit": (i, [POINTER(vp), i, i, i, i, i, POINTER(i), POINTER(i), POINTER(vp), POINTER(vp), i, i]),
        "pdmrg_heff_plan_create_host": (i, [POINTER(vp), i, i, i, i, i, POINTER(i), POINTER(i), POINTER(vp), POINTER(vp), i, i]),
        "pdmrg_heff_plan_blob_bytes": (sz, [vp]),
        "pdmrg_heff_plan_upload": (i, [vp, vp, sz, vp]),
        "pdmrg_heff_plan_destroy": (None, [vp]),
        "pdmrg_heff_workspace_bytes": (sz, [vp]),
        "pdmrg_heff_apply": (i, [vp, POINTER(vp), POINTER(vp), vp, vp, d, vp, sz, vp]),
        "pdmrg_heff_apply_timed": (i, [vp, POINTER(vp), POINTER(vp), vp, vp, d, vp, sz, vp, POINTER(ctypes.c_float)]),
        "pdmrg_ipc_alloc": (i, [sz, POINTER(vp), ctypes.c_char_p]),
        "pdmrg_ipc_open": (i, [ctypes.c_char_p, POINTER(vp)]),
        "pdmrg_ipc_close": (i, [vp]),
        "pdmrg_ipc_free": (i, [vp]),
        "pdmrg_comm_create": (i, [POINTER(vp), i, i, POINTER(vp), POINTER(vp), POINTER(vp)]),
        "pdmrg_comm_destroy": (None, [vp]),
        "pdmrg_comm_error": (i, [vp, POINTER(ctypes.c_ulonglong), vp]),
        "pdmrg_comm_z_bytes": (sz, [vp, i]),
        "pdmrg_comm_y_bytes": (sz, [vp]),
        "pdmrg_heff_needed_j": (i, [vp, i, POINTER(ctypes.c_uint)]),
        "pdmrg_heff_apply_fused": (i, [vp, vp, POINTER(vp), POINTER(vp), vp, vp, d, POINTER(ctypes.c_uint), vp, sz, vp]),
        "pdmrg_gemm_nt_scatter": (i, [i, i, i, i, i, i, POINTER(i), vp, ll, ll, ll, vp, ll, ll, ll, i, POINTER(vp), i, i, ll, ll, ll, d, d, i, vp]),
        "pdmrg_heff_apply_fused_ket": (i, [vp, vp, POINTER(vp), POINTER(vp), vp, vp, d, vp, sz, vp]),
        "pdmrg_heff_flops_dense": (d, [vp]),
        "pdmrg_heff_flops_sparse": (d, [vp]),
        "pdmrg_heff_flops_gemm": (d, [vp]),
        "pdmrg_env_workspace_bytes": (sz, [i, i, i, i, i, i]),
        "pdmrg_env_update_right": (i, [i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, sz, vp]),
        "pdmrg_env_update_left": (i, [i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, sz, vp]),
        "pdmrg_env_update_right_cols": (i, [i, i, i, i, i, i, vp, vp, vp, vp, i, i, vp, vp, sz, vp]),
        "pdmrg_env_update_left_cols": (i, [i, i, i, i, i, i, vp, vp, vp, vp, i, i, vp, vp, sz, vp]),
        "pdmrg_vec_dots": (i, [i, ll, i, vp, ll, vp, vp, vp, sz, vp]),
        "pdmrg_vec_combine": (i, [i, ll, i, vp, vp, ll, vp, i, vp]),
        "pdmrg_davidson_residual": (i, [i, ll, i, vp, vp, vp, vp, ll, vp, vp, vp, sz, vp]),
        "pdmrg_vec_project": (i, [i, ll, i, vp, ll, vp, vp, vp, vp, vp, sz, vp]),
        "pdmrg_vec_normalize": (i, [i, ll, vp, vp, vp, vp]),
        "pdmrg_vec_scratch_bytes": (sz, [i]),
        "pdmrg_davidson_scratch_bytes": (sz, [i, i]),
        "pdmrg_davidson": (i, [vp, POINTER(vp), POINTER(vp), d, i, ll, vp, ll, i, i, i, d, d, i, i, vp, vp, vp, ll, vp, sz, vp, sz,
                               POINTER(d), vp, ll, POINTER(i), POINTER(i), POINTER(d), vp]),
        "pdmrg_davidson_precond": (i, [vp, POINTER(vp), POINTER(vp), d, i, ll, vp, ll, i, i, i, d, d, i, i, vp, vp, vp, ll, vp, sz,
                                       vp, sz, POINTER(d), vp, ll, POINTER(i), POINTER(i), POINTER(d), vp, vp, d]),
        "pdmrg_davidson_ex": (i, [vp, POINTER(vp), POINTER(vp), d, i, ll, vp, ll, i, i, i, d, d, i, i, vp, vp, vp, ll, vp, sz,
                                  vp, sz, POINTER(d), vp, ll, POINTER(i), POINTER(i), POINTER(d), vp, vp, d, i, i, POINTER(i), i]),
        "pdmrg_davidson_sharded": (i, [vp, vp, POINTER(vp), POINTER(vp), d, i, ll, vp, ll, i, i, i, d, d, i, i, vp, vp, vp, ll, vp, sz,
                                       vp, sz, POINTER(d), vp, ll, POINTER(i), POINTER(i), POINTER(d), vp, vp, d, i, i, POINTER(i), i]),
        "pdmrg_vec_precond": (i, [i, ll, vp, vp, d, d, d, vp, vp, sz, vp]),
        "pdmrg_heff_diag": (i, [i, i, i, i, i, i, vp, vp, vp, d, i, vp, vp]),
        "pdmrg_host_eig": (i, [i, POINTER(d), POINTER(d), POINTER(d)]),
        "pdmrg_svd_workspace_bytes": (sz, [i, i, i, i]),
        "pdmrg_svd": (i, [i, i, i, vp, vp, vp, vp, i, vp, sz, vp, vp]),
        "pdmrg_svd_topk": (i, [i, i, i, i, vp, vp, vp, vp, i, vp, sz, vp, vp]),
        "pdmrg_left_basis_workspace_bytes": (sz, [i, i, i]),
        "pdmrg_left_basis": (i, [i, i, i, i, vp, vp, vp, vp, sz, vp, vp]),
        "pdmrg_recycle_basis_workspace_bytes": (sz, [i, i, i, i, i, i]),
        "pdmrg_recycle_basis": (i, [i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]),
        "pdmrg_heevd_workspace_bytes": (sz, [i, i]),
        "pdmrg_heevd": (i, [i, i, vp, vp, vp, sz, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.pdmrg_last_error()
        raise PdmrgError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(lib.pdmrg_launch_count())
