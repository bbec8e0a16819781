"""Golden fixture of the reference's ``two_loops`` (ADMM.py:410-508), generated from the UNTOUCHED reference.

    python tests/golden/make_golden_two_loops.py

The method returns nothing, so the iterates its locals hold at the end are captured by wrapping bound methods on the
instance (``CG_solver`` gives x, z_u, z_d of every inner iteration; ``phi_direct`` gives phi and the gamma it was called
with) - the reference source is not edited."""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refload import load_reference  # noqa: E402
from make_golden import _np, _stack_coeffs  # noqa: E402
from mixed_graph_admm_b200 import synth  # noqa: E402


def make(name, dtype, limits):
    _, ref_admm = load_reference()
    N, k, T, t_in, B = 24, 4, 6, 3, 3
    gi = synth.road_graph(N, 1.3, seed=1, isolate_pair=True)
    ai = dict(synth.admm_info(N), mu_d1=0.05)        # small DGTV weight: the soft threshold is active
    ctor = dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T)
    y = synth.signals(B, t_in, N, seed=12, dtype=dtype)
    with contextlib.redirect_stdout(io.StringIO()):
        blk = ref_admm.ADMM_algorithm(gi, ai, **ctor)
    for key, v in limits.items():
        setattr(blk, key, v)
    solves, proxes = [], []
    cg_orig, phi_orig = blk.CG_solver, blk.phi_direct

    def cg_wrap(fn, rhs, x0=None, **kw):
        out = cg_orig(fn, rhs, x0, **kw)
        solves.append(out[0])
        return out

    def phi_wrap(x, gamma):
        out = phi_orig(x, gamma)
        proxes.append((x, gamma, out))
        return out

    blk.CG_solver, blk.phi_direct = cg_wrap, phi_wrap
    with contextlib.redirect_stdout(io.StringIO()):
        ret = blk.two_loops(y)
    assert ret is None
    x, zu, zd = solves[-3], solves[-2], solves[-1]
    xl, gl, phil = proxes[-1]
    assert torch.equal(xl, x)
    # gamma_u / gamma_d of the last outer iteration: reset to 0.1, then one ascent per inner iteration
    n_in = limits["max_inner_iter"]
    gu = torch.ones_like(x) * 0.1
    gd = torch.ones_like(x) * 0.1
    for i in range(n_in):
        xs, zus, zds = solves[-3 * (n_in - i)], solves[-3 * (n_in - i) + 1], solves[-3 * (n_in - i) + 2]
        gu = gu + blk.rho_u * (xs - zus)
        gd = gd + blk.rho_d * (xs - zds)
    d = {"meta": json.dumps({"ctor": ctor, "limits": limits, "admm_info": ai, "init": None,
                             "dtype": str(dtype).replace("torch.", ""), "n_nodes": N}),
         "u_edges": _np(gi["u_edges"]), "u_dist": _np(gi["u_dist"]), "y": _np(y),
         "connect_list": _np(blk.connect_list), "dist_list": _np(blk.dist_list), "u_ew": _np(blk.u_ew), "d_ew": _np(blk.d_ew),
         "x": _np(x), "zu": _np(zu), "zd": _np(zd), "phi": _np(phil), "gamma": _np(gl + blk.rho * (phil - blk.apply_op_Ldr(xl))),
         "gamma_u": _np(gu), "gamma_d": _np(gd),
         "cg_iter_x": np.array(blk.CG_iter_x), "cg_iter_zu": np.array(blk.CG_iter_zu), "cg_iter_zd": np.array(blk.CG_iter_zd),
         "alpha_x": _stack_coeffs(blk.alpha_x), "beta_x": _stack_coeffs(blk.beta_x),
         "n_lists": np.array([len(blk.p_res_list), len(blk.x_shift_list), len(blk.GLR_list)])}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB, {len(blk.CG_iter_x)} x-solves, x[0,:3,0,0]={d['x'][0, :3, 0, 0]}")


if __name__ == "__main__":
    make("two_loops_f32", torch.float32, {"max_ADMM_iter": 3, "max_inner_iter": 2, "max_CG_iter": 5, "CG_tol": -1.0})
    make("two_loops_f64", torch.float64, {"max_ADMM_iter": 2, "max_inner_iter": 3, "max_CG_iter": 6, "CG_tol": -1.0})
