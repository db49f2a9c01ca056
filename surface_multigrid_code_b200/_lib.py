"""ctypes binding of ``libsmg.so`` (the C ABI declared in ``include/smg.h``).

There is deliberately no fallback: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
surface_multigrid_code_b200/csrc``) loading raises, and without a CUDA device
``smg_create`` fails with ``SMG_E_CUDA`` (except for plan-only handles, which can
only do the host-side index planning).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMG_LIB_PATH: load another build of the same library (profiling experiments only)
LIB_PATH = os.environ.get("SMG_LIB_PATH") or os.path.join(_HERE, "libsmg.so")

SMG_DEVICE_CURRENT = -1
SMG_DEVICE_NONE = -2

SMG_OK = 0
STATUS = {
    0: "SMG_OK",
    1: "SMG_E_INVALID",
    2: "SMG_E_CUDA",
    3: "SMG_E_NLEVELS",
    4: "SMG_E_NONFINITE",
    5: "SMG_E_STATE",
    6: "SMG_E_CUSOLVER",
    7: "SMG_E_NCCL",
    8: "SMG_E_NOT_SYMMETRIC",
    9: "SMG_E_UNSUPPORTED",
    10: "SMG_E_INTERNAL",
}

SMOOTHER_WAVEFRONT = 0
SMOOTHER_MULTICOLOUR = 1

MAT = {"A": 0, "P": 1, "PT": 2, "LHS": 3, "Auk": 4}
KERNEL = {
    "residual": 0,
    "relax_sweep": 1,
    "restrict": 2,
    "prolong_add": 3,
    "residual_norm": 4,
    "coarse_solve": 5,
    "vcycle": 6,
    "mg_iteration": 7,
    "relax_pre": 8,
}


class smg_options(C.Structure):
    _fields_ = [
        ("pre_relax", C.c_int),
        ("post_relax", C.c_int),
        ("smoother", C.c_int),
        ("device", C.c_int),
        ("use_graph", C.c_int),
        ("verbose", C.c_int),
        ("locality_reorder", C.c_int),
        ("sigma", C.c_int),
        ("patch_rows", C.c_int),
        ("reserved0", C.c_int),
        ("reserved", C.c_int * 6),
    ]


_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/smg.h declares
SIGNATURES = {
    "smg_default_options": (None, [C.POINTER(smg_options)]),
    "smg_version": (C.c_int, []),
    "smg_status_string": (C.c_char_p, [C.c_int]),
    "smg_last_error": (C.c_char_p, [_vp]),
    "smg_create": (C.c_int, [C.POINTER(_vp), C.POINTER(smg_options)]),
    "smg_destroy": (None, [_vp]),
    "smg_set_hierarchy": (C.c_int, [_vp, C.c_int, _ip, C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_dp)]),
    "smg_precompute": (C.c_int, [_vp, C.c_int, _ip, _ip, _dp, _ip, C.c_int]),
    "smg_update_values": (C.c_int, [_vp, _dp]),
    "smg_solve": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int, _dp, _dp, _ip, _ip]),
    "smg_solve_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_double, C.c_int, _vp, _dp, _ip, _ip]),
    "smg_mcf_setup": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _dp, C.c_double]),
    "smg_mcf_step": (C.c_int, [_vp, _dp, C.c_double, C.c_int, _dp, _dp, _ip, _ip]),
    "smg_vcycle": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "smg_relax": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "smg_apply_A": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "smg_residual": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int]),
    "smg_residual_norm": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int, _dp]),
    "smg_restrict": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "smg_prolong": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "smg_coarse_solve": (C.c_int, [_vp, _dp, _dp, C.c_int]),
    "smg_dist_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_size_t]),
    "smg_dist_handle_bytes": (C.c_int, []),
    "smg_dist_get_handle": (C.c_int, [_vp, _vp]),
    "smg_dist_connect": (C.c_int, [_vp, _vp]),
    "smg_rendezvous_files": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, _vp, C.c_size_t, _vp, C.c_int]),
    "smg_dist_connect_files": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_int]),
    "smg_dist_set_options": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "smg_dist_info": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "smg_dist_level_info": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "smg_dist_get_part": (C.c_int, [_vp, C.c_int, _ip]),
    "smg_dist_get_exchange": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip]),
    "smg_num_levels": (C.c_int, [_vp]),
    "smg_level_rows": (C.c_int, [_vp, C.c_int]),
    "smg_num_unknown": (C.c_int, [_vp]),
    "smg_get_unknown": (C.c_int, [_vp, _ip]),
    "smg_get_keep": (C.c_int, [_vp, C.c_int, _ip, _ip]),
    "smg_matrix_dims": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _ip]),
    "smg_matrix_copy": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _dp]),
    "smg_get_diag": (C.c_int, [_vp, C.c_int, _dp]),
    "smg_get_phases": (C.c_int, [_vp, C.c_int, _ip, _ip]),
    "smg_get_row_order": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip]),
    "smg_level_padded_nnz": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "smg_level_stats": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "smg_patch_plan": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "smg_level_patched": (C.c_int, [_vp, C.c_int]),
    "smg_solve_on_device": (C.c_int, [_vp]),
    "smg_time_kernel": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), _ip]),
    "smg_trace_iteration": (C.c_int, [_vp, C.c_int, C.c_int, C.c_char_p, C.c_int, _dp, _dp, _ip]),
    "smg_launch_count": (C.c_int64, [_vp]),
    "smg_get_timings": (C.c_int, [_vp, _dp, C.c_int]),
    "smg_get_stream": (_vp, [_vp]),
}

_lib = None


def load():
    """Load libsmg.so and declare every prototype. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(make -C surface_multigrid_code_b200/csrc, or __graft_entry__.build()); "
            "this package has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
