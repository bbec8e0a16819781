"""Import the untouched reference from /root/reference (authoring container only).

``/root/reference`` does not exist on the GPU box, so nothing marked ``gpu`` may use this;
it serves ``tests/golden/make_golden.py`` and the live oracle-vs-reference checks in
``tests/test_oracle.py`` (skipped when the directory is absent).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import torch

REF_DIR = os.environ.get("MGA_REFERENCE_DIR", "/root/reference")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "ADMM.py"))


def load_reference():
    """Returns ``(utils_module, ADMM_module)`` of the reference, or ``None``."""
    if not have_reference():
        return None
    if "matplotlib" not in sys.modules:        # ADMM.py:7 imports pyplot; it is not installed
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            mp = types.ModuleType("matplotlib.pyplot")
            m.pyplot = mp
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = mp
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils as ref_utils   # the reference's, found through REF_DIR
        import ADMM as ref_admm
    assert os.path.dirname(os.path.abspath(ref_admm.__file__)) == os.path.abspath(REF_DIR)
    return ref_utils, ref_admm


def run_reference(graph_info, admm_info, y, ctor, limits, mask=None, init=None):
    """Run the reference's ``combined_loop`` and capture the iterates it keeps local.

    ``ctor``: kwargs of ``ADMM_algorithm``; ``limits``: dict of the mutable attributes
    ``max_ADMM_iter, max_CG_iter, CG_tol, ADMM_tol``; ``init``: optional
    ``(ablation, use_line_graph)`` passed to ``init_iterations`` before the run.
    z_u, z_d, phi and the duals are recovered by wrapping bound methods on the instance
     — the reference source is not edited.
    """
    _, ref_admm = load_reference()
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        blk = ref_admm.ADMM_algorithm(graph_info, admm_info, **ctor)
        if init is not None:
            blk.init_iterations(*init)
    for k, v in limits.items():
        setattr(blk, k, v)
    solves, proxes = [], []
    cg_orig, phi_orig = blk.CG_solver, blk.phi_direct

    def cg_wrap(fn, rhs, x0=None, **kw):
        out = cg_orig(fn, rhs, x0, **kw)
        solves.append((rhs, out[0]))
        return out

    def phi_wrap(x, gamma):
        out = phi_orig(x, gamma)
        proxes.append((x, gamma, out))
        return out

    blk.CG_solver = cg_wrap
    blk.phi_direct = phi_wrap
    with contextlib.redirect_stdout(sink):
        x = blk.combined_loop(y, mask=mask, print_info=False)
    per_outer = 2 if blk.ablation == 'DGLR' else 3
    n_outer = len(blk.CG_iter_x)
    xs = [solves[per_outer * i][1] for i in range(n_outer)]
    zus = [solves[per_outer * i + 1][1] for i in range(n_outer)]
    x0 = ref_admm.initial_guess(y, blk.t_in, blk.T) if mask is None else None
    if mask is not None:
        with contextlib.redirect_stdout(sink):
            x0 = ref_admm.initial_interpolation(y, mask)
    gu = torch.ones_like(x0) * 0.1
    gd = torch.ones_like(x0) * 0.1
    for i in range(n_outer):
        gu = gu + blk.rho_u * (xs[i] - zus[i])
        if per_outer == 3:
            gd = gd + blk.rho_d * (xs[i] - solves[3 * i + 2][1])
    out = {"x": x, "zu": zus[-1], "gamma_u": gu, "gamma_d": gd}
    if per_outer == 3:
        out["zd"] = solves[3 * n_outer - 1][1]
    if proxes:
        xl, gl, phil = proxes[-1]
        out["phi"] = phil
        out["gamma"] = gl + blk.rho * (phil - blk.apply_op_Ldr(xl))
    out["blk"] = blk
    return out
