"""ctypes binding of libcrb.so (C ABI declared in include/crb.h).

There is no CPU fallback: if the shared library is missing or was built for another ABI
version, importing a symbol raises ``RuntimeError`` naming the build command.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

CRB_MAX_SLOTS = 256
CRB_VERSION = 107

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRB_LIB", os.path.join(_HERE, "libcrb.so"))  # CRB_LIB: kernel-variant experiments


class CrbPlan(C.Structure):
    _fields_ = [
        ("n_elements", C.c_int32),
        ("n_free", C.c_int32),
        ("n0", C.c_int32),
        ("p_act", C.c_int32),
        ("m", C.c_int32),
        ("g", C.c_int32),
        ("p", C.c_int32),
        ("levels", C.c_int32),
        ("contiguous", C.c_int32),
        ("has_mask", C.c_int32),
        ("mfac_doubles", C.c_int64),
        ("kcoef_doubles", C.c_int64),
        ("red_index", C.c_int32 * (3 * CRB_MAX_SLOTS)),
    ]


class CrbSystem(C.Structure):
    _fields_ = [
        ("n_members", C.c_int32),
        ("mass_shared", C.c_int32),
        ("stiff_shared", C.c_int32),
        ("force_shared", C.c_int32),
        ("mfac", C.c_void_p),
        ("kcoef", C.c_void_p),
        ("elem_type", C.c_void_p),
        ("red_index", C.c_void_p),
        ("drag", C.c_void_p),
        ("grav", C.c_void_p),
        ("seg_half_mass", C.c_void_p),
        ("gx", C.c_double),
        ("gy", C.c_double),
        ("grav_mode", C.c_int32),
        ("u_const", C.c_void_p),
        ("imp_amp", C.c_void_p),
        ("imp_dof", C.c_int32),
        ("imp_duration", C.c_double),
        ("gain", C.c_void_p),
        ("ref", C.c_void_p),
        ("gain_frag", C.c_void_p),
        ("f_ext", C.c_void_p),
        ("all_linear", C.c_int32),
        ("all_nonlinear", C.c_int32),
        ("force_general", C.c_int32),
        ("force_staged", C.c_int32),
        ("shared_op", C.c_void_p),
        ("shared_op_doubles", C.c_int64),
        ("gain_stride", C.c_int64),
        ("member_op", C.c_void_p),
        ("u_sin_amp", C.c_void_p),
        ("u_sin_omega", C.c_double),
        ("u_sin_phase", C.c_double),
        ("u_tab_t", C.c_void_p),
        ("u_tab_v", C.c_void_p),
        ("u_tab_k", C.c_int32),
        ("u_time_shared", C.c_int32),
        ("tile_counter", C.c_void_p),
        ("out_sel_inv", C.c_void_p),
        ("out_n_sel", C.c_int32),
        ("member_order", C.c_void_p),
    ]


_SIGNATURES = {
    "crb_version": (C.c_int, []),
    "crb_abi_sizes": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crb_probe_dfma": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "crb_last_error": (C.c_char_p, []),
    "crb_plan": (C.c_int, [C.c_int32, C.c_char_p, C.c_int32, C.POINTER(CrbPlan)]),
    "crb_assemble": (
        C.c_int,
        [C.POINTER(CrbPlan), C.c_void_p, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32,
         C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "crb_rhs": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_double,
                          C.c_void_p, C.c_void_p]),
    "crb_forces": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_void_p, C.c_void_p]),
    "crb_rk4": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_double,
                          C.c_double, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "crb_rk45": (
        C.c_int,
        [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
         C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_int32, C.c_void_p],
    ),
    "crb_rk45_stage": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "crb_rk45_control": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "crb_assemble_shifted": (C.c_int, [C.POINTER(CrbPlan), C.c_void_p, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32,
                                       C.c_double, C.c_void_p, C.c_void_p]),
    "crb_midpoint": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_int32, C.c_void_p, C.c_double,
                               C.c_double, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "crb_rk4_wave_members": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.POINTER(C.c_int32)]),
    "crb_system_slice": (C.c_int, [C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_int32, C.c_int32, C.POINTER(CrbSystem)]),
    "crb_pipeline_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "crb_pipeline_destroy": (C.c_int, [C.c_void_p]),
    "crb_pipeline_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crb_pipeline_synchronize": (C.c_int, [C.c_void_p]),
    "crb_rk4_host": (C.c_int, [C.c_void_p, C.POINTER(CrbPlan), C.POINTER(CrbSystem), C.c_void_p, C.c_void_p,
                               C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p]),
    "crb_gain_fragments": (C.c_int64, [C.POINTER(CrbPlan), C.c_void_p, C.c_void_p]),
    "crb_shared_operator": (C.c_int64, [C.POINTER(CrbPlan), C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_void_p]),
    "crb_shared_sparse_masks": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p]),
    "crb_dense_matrices": (C.c_int, [C.POINTER(CrbPlan), C.c_void_p, C.c_char_p, C.c_char_p,
                                     C.c_void_p, C.c_void_p]),
    "crb_dense_matrices_batched": (C.c_int, [C.POINTER(CrbPlan), C.c_void_p, C.c_int32, C.c_char_p, C.c_char_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "crb_lqr_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "crb_member_operators": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crb_lqr_gains": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libcrb.so once; fail loudly if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` from the repo root."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.crb_version()
    if v != CRB_VERSION:
        raise RuntimeError(f"libcrb.so reports ABI version {v}, expected {CRB_VERSION}: rebuild it")
    pb, sb = C.c_int32(), C.c_int32()
    lib.crb_abi_sizes(C.byref(pb), C.byref(sb))
    if (pb.value, sb.value) != (C.sizeof(CrbPlan), C.sizeof(CrbSystem)):
        raise RuntimeError(f"struct layout mismatch: libcrb.so has crb_plan_t {pb.value} B / crb_system_t {sb.value} B, "
                           f"the ctypes mirror {C.sizeof(CrbPlan)} / {C.sizeof(CrbSystem)}")
    _lib = lib
    return lib


class CrbError(RuntimeError):
    pass


def check(rc: int, exc=None) -> None:
    """Map a negative return code to a Python exception carrying crb_last_error()."""
    if rc == 0:
        return
    msg = load().crb_last_error().decode("utf-8", "replace")
    if exc is None:
        exc = ValueError if rc in (-1, -3) else CrbError
    raise exc(msg)
