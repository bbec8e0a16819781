"""ctypes binding of libmga.so — the only way the Python host code reaches the kernels.

The signatures mirror ``include/mga.h`` one to one.  There is no fallback: if the library
is missing, ``lib()`` raises, and so does every solver entry point built on it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# MGA_LIB selects an experiment build of the same library (mixed_graph_admm_b200/build.py --tag=...)
LIB_PATH = os.environ.get("MGA_LIB") or os.path.join(HERE, "_lib", "libmga.so")

MGA_F32, MGA_F64 = 0, 1
LDRT_SCATTER, LDRT_GATHER = 0, 1
TEMPORAL_GRAPH, TEMPORAL_LINE, TEMPORAL_BAND = 0, 1, 2
ABLATION = {"None": 0, "DGTV": 1, "DGLR": 2, "UT": 3}
OP = {"Lu": 0, "Ldr": 1, "Ldr_T": 2, "cLdr": 3, "LHS_x": 4, "LHS_zu": 5, "LHS_zd": 6}
SYS = {"x": 0, "zu": 1, "zd": 2}
MODE = {"auto": 0, "streaming": 1, "resident": 2, "streaming_point": 3}
DIAG_COLS = 12
(DIAG_DX2, DIAG_X_ZU2, DIAG_DZU2, DIAG_GLR, DIAG_RECOVER2, DIAG_PHI_LDX2, DIAG_DPHI2, DIAG_DGTV, DIAG_X_ZD2,
 DIAG_DZD2, DIAG_DGLR, DIAG_NONFINITE) = range(12)
ERR_INVALID, ERR_INDEX, ERR_CUDA, ERR_UNSUPPORTED, ERR_NONFINITE = -1, -2, -3, -4, -5

EXPORTS = [
    "mga_plan_create", "mga_plan_destroy", "mga_plan_resident_eligible", "mga_plan_info", "mga_plan_set_cg_mode", "mga_apply",
    "mga_cg_solve", "mga_initial_guess", "mga_rhs_x", "mga_dual_ascent", "mga_prox_phi_dual", "mga_phi_direct",
    "mga_admm_solve", "mga_admm_solve_host", "mga_cluster_solve", "mga_knn_build", "mga_schedule_selfcheck", "mga_launch_count", "mga_last_error", "mga_version",
]


class GraphDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("T", C.c_int32), ("t_in", C.c_int32),
                ("ku", C.c_int32), ("nbr_u", C.c_void_p), ("u_w", C.c_void_p), ("u_w_T", C.c_int32),
                ("kd", C.c_int32), ("nbr_d", C.c_void_p), ("d_w", C.c_void_p), ("d_w_T", C.c_int32),
                ("ldrt_mode", C.c_int32), ("temporal", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("rho", C.c_double), ("rho_u", C.c_double), ("rho_d", C.c_double), ("mu_u", C.c_double),
                ("mu_d1", C.c_double), ("mu_d2", C.c_double), ("ablation", C.c_int32), ("reserved", C.c_int32)]


class AdmmOutputs(C.Structure):
    _fields_ = [("zu", C.c_void_p), ("zd", C.c_void_p), ("phi", C.c_void_p), ("gamma", C.c_void_p),
                ("gamma_u", C.c_void_p), ("gamma_d", C.c_void_p), ("diag", C.c_void_p), ("dx_sum", C.c_void_p),
                ("alpha", C.c_void_p), ("beta", C.c_void_p), ("cg_iters", C.c_void_p), ("outer_done", C.c_void_p)]


_lib = None
_lock = threading.Lock()


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    """The loaded library.  Raises if it was never built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m mixed_graph_admm_b200.build` "
                "(needs nvcc). The solver has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.mga_version.restype = C.c_int
        L.mga_last_error.restype = C.c_char_p
        L.mga_launch_count.restype = C.c_int64
        L.mga_plan_create.argtypes = [C.POINTER(GraphDesc), C.c_int, C.POINTER(vp)]
        L.mga_plan_destroy.argtypes = [vp]
        L.mga_plan_destroy.restype = None
        L.mga_plan_resident_eligible.argtypes = [vp, C.c_int]
        L.mga_plan_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.mga_plan_set_cg_mode.argtypes = [vp, C.c_int]
        L.mga_apply.argtypes = [vp, C.c_int, C.POINTER(Params), vp, vp, vp, i64, C.c_int, vp]
        L.mga_cg_solve.argtypes = [vp, C.c_int, C.POINTER(Params), vp, vp, vp, i64, C.c_int, C.c_int, dbl,
                                   C.POINTER(i32), vp, vp, vp]
        L.mga_initial_guess.argtypes = [vp, vp, vp, i64, C.c_int, dbl, dbl, vp]
        L.mga_rhs_x.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, i64, C.c_int, vp]
        L.mga_dual_ascent.argtypes = [vp, dbl, vp, vp, vp, i64, C.c_int, vp]
        L.mga_prox_phi_dual.argtypes = [vp, C.POINTER(Params), vp, vp, vp, i64, C.c_int, vp]
        L.mga_phi_direct.argtypes = [vp, C.POINTER(Params), vp, vp, vp, i64, C.c_int, vp]
        L.mga_admm_solve.argtypes = [vp, C.POINTER(Params), vp, C.c_int, vp, vp, i64, C.c_int, C.c_int, C.c_int,
                                     dbl, dbl, dbl, dbl, C.c_int, C.POINTER(AdmmOutputs), C.c_int, vp]
        L.mga_cluster_solve.argtypes = [vp, C.POINTER(Params), vp, vp, i64, C.c_int, C.c_int, C.c_int, dbl, dbl, dbl, dbl,
                                        C.c_int, C.POINTER(AdmmOutputs), vp]
        L.mga_admm_solve_host.argtypes = [vp, C.POINTER(Params), vp, C.c_int, vp, i64, C.c_int, C.c_int, C.c_int,
                                          dbl, dbl, C.c_int, vp, vp, vp, vp, C.c_int, i64]
        L.mga_knn_build.argtypes = [i32, i64, vp, vp, i32, vp, vp]
        L.mga_schedule_selfcheck.argtypes = [C.POINTER(GraphDesc), C.POINTER(dbl)]
        for name in EXPORTS:
            getattr(L, name)        # AttributeError here = header and library out of step
        _lib = L
        return _lib


class MgaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libmga error {code}: {text}")
        self.code = code
        self.text = text


def check(rc: int):
    if rc == 0:
        return
    text = lib().mga_last_error().decode("utf-8", "replace")
    if rc == ERR_INDEX:
        raise ValueError(text or "Index out of bounds")       # ADMM.py:205-206
    raise MgaError(rc, text)


def dtype_id(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return MGA_F32
    if dt == torch.float64:
        return MGA_F64
    raise TypeError(f"libmga computes in float32 or float64, got {dt}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def knn_build(n_nodes, edges, dists, k):
    """``mga_knn_build``: (N, k+1) int32 / float32 tables, bit-identical to utils.py:183-204."""
    L = lib()
    e = torch.as_tensor(edges).to(torch.int64).contiguous()
    d = torch.as_tensor(dists).to(torch.float64).contiguous()
    nodes = torch.empty((n_nodes, k + 1), dtype=torch.int32)
    nd = torch.empty((n_nodes, k + 1), dtype=torch.float32)
    rc = L.mga_knn_build(int(n_nodes), int(e.shape[0]), ptr(e), ptr(d), int(k), ptr(nodes), ptr(nd))
    if rc == ERR_INDEX:
        raise KeyError(L.mga_last_error().decode())
    check(rc)
    return nodes, nd
