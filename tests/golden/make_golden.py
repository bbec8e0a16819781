"""Generate the golden fixtures in this directory from the UNTOUCHED reference.

Run in the authoring container only (``/root/reference`` must exist):

    python tests/golden/make_golden.py

Every fixture is an ``.npz`` with the inputs (edge list, distances, ADMM weights, ``y``,
constructor kwargs and iteration limits as a JSON string), the graph tables the reference
built (``connect_list, dist_list, u_ew, d_ew``), the outputs of ``combined_loop`` including
the iterates it keeps local (captured by wrapping bound methods, the source is not edited),
its diagnostic lists, and the reference operators applied to a seeded probe vector.
"""
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

from _refload import load_reference, run_reference  # noqa: E402
from mixed_graph_admm_b200 import synth  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _stack_coeffs(lst):
    """alpha/beta lists -> (outer, iters, B) array; ragged (tolerance mode) -> padded with NaN."""
    rows = []
    for e in lst:
        e = torch.stack(list(e)) if isinstance(e, list) else e
        rows.append(e.reshape(e.shape[0], -1).to(torch.float64))
    m = max(r.shape[0] for r in rows)
    out = torch.full((len(rows), m, rows[0].shape[1]), float('nan'), dtype=torch.float64)
    for i, r in enumerate(rows):
        out[i, :r.shape[0]] = r
    return out.numpy()


ONLY = set()     # fixture names given on the command line: regenerate just those


def make_case(name, graph_info, admm_info, y, ctor, limits, mask=None, init=None, probe_seed=123):
    if ONLY and name not in ONLY:
        return
    ref = run_reference(graph_info, admm_info, y, ctor, limits, mask=mask, init=init)
    blk = ref["blk"]
    T = blk.T
    g = torch.Generator().manual_seed(probe_seed)
    xp = torch.randn(2, T, blk.n_nodes, y.size(-1), generator=g, dtype=y.dtype)
    gp = torch.randn(2, T, blk.n_nodes, y.size(-1), generator=g, dtype=y.dtype)
    d = {
        "meta": json.dumps({"ctor": ctor, "limits": limits, "admm_info": admm_info, "init": init,
                            "dtype": str(y.dtype).replace("torch.", ""), "n_nodes": graph_info["n_nodes"]}),
        "u_edges": _np(graph_info["u_edges"]), "u_dist": _np(graph_info["u_dist"]),
        "y": _np(y),
        "connect_list": _np(blk.connect_list), "dist_list": _np(blk.dist_list),
        "u_ew": _np(blk.u_ew), "d_ew": _np(blk.d_ew),
        "probe_x": _np(xp), "probe_gamma": _np(gp),
        "op_Lu": _np(blk.apply_op_Lu(xp)), "op_Ldr": _np(blk.apply_op_Ldr(xp)),
        "op_Ldr_T": _np(blk.apply_op_Ldr_T(xp)), "op_cLdr": _np(blk.apply_op_cLdr(xp)),
        "op_LHS_x": _np(blk.LHS_x(xp)), "op_LHS_zu": _np(blk.LHS_zu(xp)),
        "op_phi_direct": _np(blk.phi_direct(xp, gp)),
        "cg_iter_x": np.array(blk.CG_iter_x), "cg_iter_zu": np.array(blk.CG_iter_zu),
        "cg_iter_zd": np.array(blk.CG_iter_zd),
        "alpha_x": _stack_coeffs(blk.alpha_x), "beta_x": _stack_coeffs(blk.beta_x),
        "alpha_zu": _stack_coeffs(blk.alpha_zu), "beta_zu": _stack_coeffs(blk.beta_zu),
        "p_res": np.array(blk.p_res_list), "d_res": np.array(blk.d_res_list),
        "x_shift": np.array(blk.x_shift_list), "recover": np.array(blk.recover_list),
        "glr": np.array([v.item() for v in blk.GLR_list]),
        "dgtv": np.array([v.item() for v in blk.DGTV_list]),
        "dglr": np.array([v.item() for v in blk.DGLR_list]),
        "delta_x_per_step": np.stack([_np(v) for v in blk.delta_x_per_step]),
    }
    if blk.ablation != 'DGLR':
        d["op_LHS_zd"] = _np(blk.LHS_zd(xp))
        d["alpha_zd"] = _stack_coeffs(blk.alpha_zd)
        d["beta_zd"] = _stack_coeffs(blk.beta_zd)
    if mask is not None:
        d["mask"] = _np(mask)
    if hasattr(blk, "time_list"):
        d["time_list"] = _np(blk.time_list)
    for k in ("x", "zu", "zd", "phi", "gamma", "gamma_u", "gamma_d"):
        if k in ref:
            d[k] = _np(ref[k])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB  x[0,:,0,0]={d['x'][0, :3, 0, 0]}")


def main():
    assert load_reference() is not None, "needs /root/reference"
    fixed = lambda o, c: {"max_ADMM_iter": o, "max_CG_iter": c, "CG_tol": -1.0, "ADMM_tol": -1.0}  # noqa: E731

    # 1. the hand-checkable 5-node anchor of SURVEY.md §8c
    e = torch.tensor([[0, 1], [1, 0], [2, 3], [3, 2], [3, 4], [4, 3]])
    gi = {"n_nodes": 5, "u_edges": e, "u_dist": torch.tensor([1., 1., 2., 2., 3., 3.])}
    ai = {"rho": 2.0, "rho_u": 3.0, "rho_d": 2.0, "mu_u": 1.0, "mu_d1": 0.5, "mu_d2": 1.0}
    x = (torch.arange(15, dtype=torch.float32) / 10).reshape(1, 3, 5, 1)
    make_case("anchor5", gi, ai, x[:, :2].contiguous(),
              dict(use_kNN=True, k=3, u_sigma=2, d_sigma=2, t_in=2, T=3), fixed(2, 3))

    # 2. tiny kNN graph with -1 padding, fp32 / fp64, fixed iterations
    N, k, T, t_in, B = 24, 4, 6, 3, 3
    gi = synth.road_graph(N, 1.3, seed=1, isolate_pair=True)
    ai = synth.admm_info(N)
    ctor = dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T)
    make_case("tiny_f32", gi, ai, synth.signals(B, t_in, N, seed=0), ctor, fixed(3, 5))
    make_case("tiny_f64", gi, ai, synth.signals(B, t_in, N, seed=0, dtype=torch.float64), ctor, fixed(3, 5))
    # 2b. small DGTV weight so that the soft-threshold is active on most entries (prox exercised)
    ai_prox = dict(ai, mu_d1=0.05)
    make_case("tiny_prox", gi, ai_prox, synth.signals(B, t_in, N, seed=10), ctor, fixed(3, 5))
    # 3. tolerance mode, B = 1: CG iteration counts (default CG_tol 1e-8), 5 outer iterations
    make_case("tiny_tol", gi, ai, synth.signals(1, t_in, N, seed=2, smooth=True), ctor, {"max_ADMM_iter": 5})
    make_case("tiny_tol_f64", gi, ai, synth.signals(1, t_in, N, seed=2, smooth=True, dtype=torch.float64), ctor,
              {"max_ADMM_iter": 5})
    # 4. default sigma (None) and time-expanded weights off
    make_case("tiny_noexpand", gi, ai, synth.signals(B, t_in, N, seed=3),
              dict(use_kNN=True, k=k, expand_time_dim=False, t_in=t_in, T=T), fixed(2, 4))
    # 5. physical adjacency (use_kNN=False: gather-form L_d^T)
    make_case("tiny_physical", gi, ai, synth.signals(B, t_in, N, seed=4),
              dict(use_kNN=False, u_sigma=50, d_sigma=50, t_in=t_in, T=T), fixed(2, 4))
    # 6. line graph, skip 1 and skip 2 (row N2)
    make_case("tiny_line1", gi, ai, synth.signals(B, t_in, N, seed=5),
              dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, use_line_graph=True), fixed(2, 4))
    make_case("tiny_line2", gi, ai, synth.signals(B, t_in, N, seed=6),
              dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, use_line_graph=True,
                   skip_connection=2), fixed(2, 4))
    # 7. init_iterations quirk Q8: d_ew reset to default sigma, un-expanded
    make_case("tiny_reinit", gi, ai, synth.signals(B, t_in, N, seed=7), ctor, fixed(2, 4), init=("None", False))
    # 8. ablations that run ('DGTV' drops the DGTV split; 'UT' same system as 'DGTV')
    make_case("tiny_abl_dgtv", gi, ai, synth.signals(B, t_in, N, seed=8),
              dict(ctor, ablation='DGTV'), fixed(2, 4))
    # 9. interpolation (mask) mode (row N3)
    # B = 1: the reference's initial_interpolation broadcasts w (B,N,C) against t (B,T,N,C)
    # (ADMM.py:802), which only lines up for B == 1
    ym = synth.signals(1, T, N, seed=9)
    mask = (torch.rand(1, T, N, 1, generator=torch.Generator().manual_seed(42)) >= 0.4).float()
    mask[:, 0] = 1
    mask[:, -1] = 1
    make_case("tiny_mask", gi, ai, ym * mask, ctor, fixed(2, 4), mask=mask)
    # 9a. the notebooks' interpolation call: float64, tolerance mode (B = 1), and float64 with fixed counts
    ym64 = synth.signals(1, T, N, seed=9, dtype=torch.float64, smooth=True)
    make_case("tiny_mask_tol_f64", gi, ai, ym64 * mask.double(), ctor, {"max_ADMM_iter": 5}, mask=mask.double())
    make_case("tiny_mask_f64", gi, ai, ym64 * mask.double(), ctor, fixed(3, 5), mask=mask.double())
    # 9b. multi-channel signals (row N4): C = 2 on the kNN graph, C = 3 on the banded line graph and on the
    # physical adjacency; same weights on every channel, dot products over (T, N, C)
    yc = lambda C, seed, dt=torch.float32: torch.rand(B, t_in, N, C, generator=torch.Generator().manual_seed(seed),  # noqa: E731
                                                         dtype=dt)
    make_case("tiny_c2", gi, ai_prox, yc(2, 20), ctor, fixed(3, 5))
    make_case("tiny_c2_f64", gi, ai, yc(2, 21, torch.float64), ctor, fixed(2, 4))
    make_case("tiny_c3_line2", gi, ai, yc(3, 22),
              dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, use_line_graph=True,
                   skip_connection=2), fixed(2, 4))
    make_case("tiny_c3_physical", gi, ai, yc(3, 23),
              dict(use_kNN=False, u_sigma=50, d_sigma=50, t_in=t_in, T=T), fixed(2, 4))
    make_case("tiny_c2_tol", gi, ai, synth.signals(1, t_in, N, seed=2, smooth=True).repeat(1, 1, 1, 2)
              * torch.tensor([1.0, 0.5]), ctor, {"max_ADMM_iter": 4})
    # 9c. mask mode with two channels (B = 1, see 9.)
    ymc = torch.rand(1, T, N, 2, generator=torch.Generator().manual_seed(30))
    maskc = (torch.rand(1, T, N, 2, generator=torch.Generator().manual_seed(31)) >= 0.4).float()
    maskc[:, 0] = 1
    maskc[:, -1] = 1
    make_case("tiny_c2_mask", gi, ai, ymc * maskc, ctor, fixed(2, 4), mask=maskc)
    # 10. PEMS08-shaped (BASELINE.json configs[0]); 4 of the 32 windows are kept
    N, k, T, t_in = 170, 6, 12, 6
    gi = synth.road_graph(N, 1.7, seed=8, isolate_pair=True)
    ctor = dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T)
    make_case("pems08_f32", gi, synth.admm_info(N), synth.signals(32, t_in, N, seed=0)[:4].contiguous(), ctor,
              fixed(5, 10))
    make_case("pems08_tol", gi, synth.admm_info(N), synth.signals(1, t_in, N, seed=1, smooth=True), ctor,
              {"max_ADMM_iter": 5})
    # 11. PEMS04-shaped graph, 2 windows (BASELINE.json configs[1] shape)
    N = 307
    gi = synth.road_graph(N, 1.1, seed=4)
    make_case("pems04_f32", gi, synth.admm_info(N), synth.signals(1024, t_in, N, seed=0)[:2].contiguous(), ctor,
              fixed(5, 10))
    # 12. PEMS04 shape in tolerance mode (B = 1, class defaults CG_tol 1e-8): the CG counts of BASELINE.md §2
    make_case("pems04_tol", gi, synth.admm_info(N), synth.signals(1, t_in, N, seed=1, smooth=True), ctor,
              {"max_ADMM_iter": 5})
    # 13. the reference's own call pattern (notebooks): B = 1, float64, T = 24, tolerances
    ctor24 = dict(use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=12, T=24)
    make_case("pems04_t24_tol_f64", gi, synth.admm_info(N), synth.signals(1, 12, N, seed=1, smooth=True, dtype=torch.float64),
              ctor24, {"max_ADMM_iter": 6})


if __name__ == "__main__":
    ONLY.update(sys.argv[1:])
    with contextlib.redirect_stderr(io.StringIO()):
        main()
