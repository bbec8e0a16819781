"""CPU oracle for the Mixed-Graph-ADMM solver hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
shipped solver (``mixed_graph_admm_b200``) never imports anything under ``oracle/`` and
raises when its CUDA library is missing.

What it is: a torch-CPU restatement of the reference's unrolled ADMM loop, function by
function, in the reference's own arithmetic order (same torch ops at the same rounding
points), so that on identical inputs it reproduces the reference bit for bit.  Each function
cites the ``/root/reference`` lines it follows.

Parity pin: the reference holds **no tests or golden vectors for this path** (SURVEY.md §4,
§8c).  The oracle is pinned instead against (1) outputs of the reference itself, imported
from ``/root/reference`` in the authoring container — ``tests/golden/make_golden.py``
generated the committed fixtures ``tests/golden/*.npz`` and ``tests/test_oracle.py``
requires ``torch.equal`` between oracle and fixture; when ``/root/reference`` is present
the same test also re-runs the reference live; (2) the reference's only operator-level
known answers, ``directed_graph.ipynb`` cells 5-12 (skip-2 line graph on ``[1..5]``) and
``CG_script.py:49-50`` (2x2 SPD system).

Conventions: signals are ``(B, T, N, C)``; ``nbr`` is the reference's ``connect_list``
``(N, K)`` int64 with the node itself in column 0 and ``-1`` padding; ``u_w`` is
``(T, N, K-1)`` or ``(N, K-1)``; ``d_w`` is ``(T-1, N, K)`` or ``(N, K)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class OracleGraph:
    """The graph tensors ``ADMM_algorithm.__init__`` leaves behind (``ADMM.py:25-52``)."""
    nbr: torch.Tensor
    u_w: torch.Tensor
    d_w: torch.Tensor
    use_knn: bool = True
    line_graph: bool = False
    skip: int = 1
    time_list: Optional[torch.Tensor] = None     # (T, skip) for line_graph with skip > 1

    @property
    def n(self):
        return self.nbr.shape[0]


@dataclass
class OracleParams:
    rho: float
    rho_u: float
    rho_d: float
    mu_u: float
    mu_d1: float
    mu_d2: float
    t_in: int = 12
    T: int = 24
    ablation: str = 'None'


def _zero_padded(x):
    # ADMM.py:143-144 — one zero column at node index N, so that index -1 reads 0
    return torch.cat((x, torch.zeros_like(x[:, :, 0:1])), 2)


def op_lu(g: OracleGraph, x):
    """``apply_op_Lu`` (ADMM.py:138-148): x - sum_j u_w[t,i,j] * x[t, nbr[i,j+1]]."""
    B, T, C = x.size(0), x.size(1), x.size(-1)
    px = _zero_padded(x)
    picked = px[:, :, g.nbr[:, 1:].reshape(-1)].reshape(B, T, g.n, -1, C)
    return x - (g.u_w.unsqueeze(0).unsqueeze(-1) * picked).sum(3)


def op_ldr(g: OracleGraph, x):
    """``apply_op_Ldr`` (ADMM.py:150-177).  Row t=0 is zero; row t>=1 is
    x[t] - sum_j d_w[t-1,i,j] * x[t-1, nbr[i,j]]."""
    B, T, N, C = x.size(0), x.size(1), x.size(2), x.size(-1)
    if g.line_graph:
        if g.skip == 1:                                   # ADMM.py:153-157
            y = x.clone()
            y[:, 0] = x[:, 0] * 0
            y[:, 1:] = x[:, 1:] - x[:, :-1]
            return y
        # ADMM.py:158-164 — banded temporal stencil, weights (T, skip, N)
        feat = (g.d_w.view(-1, N)[None, :, :, None] * x[:, g.time_list.view(-1), :, :]).reshape(B, T, -1, N, C)
        assert torch.all(feat[:, 0] == 0), "Features at time 0 should be all zero"
        y = x - feat.sum(2)
        y[:, 0] = y[:, 0] * 0
        return y
    px = _zero_padded(x)
    child = g.d_w.unsqueeze(0).unsqueeze(-1) * px[:, :-1, g.nbr.view(-1)].view(B, T - 1, g.n, -1, C)
    y = x.clone()
    y[:, 1:] = x[:, 1:] - child.sum(3)
    y[:, 0] = x[:, 0] * 0
    return y


def op_ldr_t(g: OracleGraph, x):
    """``apply_op_Ldr_T`` (ADMM.py:179-223).

    kNN branch: f[t, c] = sum over (i, j) with nbr[i,j] == c of d_w[t,i,j] * x[t+1, i], by
    scatter_add into an N+1 slot buffer (slot N swallows the -1 entries).  Result rows:
    t <= T-2: x[t] - f[t]; t = T-1: x[T-1].  Quirk Q1: line 221 zeroes row 0 but line 222
    overwrites it from x, so row 0 is x[0] - f[0] (not -f[0])."""
    B, T, C = x.size(0), x.size(1), x.size(-1)
    if g.line_graph:
        if g.skip == 1:                                   # ADMM.py:182-186
            y = x.clone()
            y[:, 0] = x[:, 0] * 0
            y[:, :-1] = y[:, :-1] - x[:, 1:]
            return y
        feat = g.d_w[None, :, :, :, None] * x[:, :, None, :, :]          # ADMM.py:188-194
        feat = torch.stack([feat.diagonal(offset=-o, dim1=1, dim2=2).sum(-1) for o in range(1, T)], dim=1)
        y = x.clone()
        y[:, 0] = x[:, 0] * 0
        y[:, :-1] = y[:, :-1] - feat
        return y
    if g.use_knn:
        held = g.d_w.unsqueeze(0).unsqueeze(-1) * x[:, 1:].unsqueeze(3)
        f = torch.zeros((B, T - 1, g.n + 1, C), dtype=held.dtype)
        idx = g.nbr.reshape(-1)[None, None, :, None].repeat(B, T - 1, 1, C)
        idx[idx == -1] = g.n
        if torch.any(idx < 0) or torch.any(idx >= f.size(2)):
            raise ValueError("Index out of bounds")
        f = f.scatter_add(2, idx, held.view(B, T - 1, -1, C))[:, :, :-1]
    else:
        # physical adjacency: a gather with the forward table (ADMM.py:211-215)
        px = _zero_padded(x)
        f = (g.d_w.unsqueeze(0).unsqueeze(-1) * px[:, 1:, g.nbr.view(-1)].view(B, T - 1, g.n, -1, C)).sum(3)
    y = x.clone()
    y[:, 0] = x[:, 0] * 0
    y[:, :-1] = x[:, :-1] - f
    return y


def op_cldr(g, x):
    """``apply_op_cLdr`` (ADMM.py:225-228)."""
    return op_ldr_t(g, op_ldr(g, x))


def glr(g, x):
    """ADMM.py:245-246."""
    return (x * op_lu(g, x)).sum((1, 2, 3)).mean()


def dglr(g, x):
    """ADMM.py:230-235."""
    return (op_ldr(g, x) ** 2).sum((1, 2, 3)).mean()


def dgtv(g, x):
    """ADMM.py:238-243."""
    return op_ldr(g, x).norm(dim=[1, 2, 3], p=1).mean()


def lhs_x(g, prm: OracleParams, x, mask=None):
    """``LHS_x`` (ADMM.py:371-387), all four ablation settings."""
    hx = x.clone()
    if mask is None:
        hx[:, prm.t_in:] = hx[:, prm.t_in:] * 0
    else:
        hx = x * mask
    if prm.ablation == 'None':
        return hx + (prm.rho_u + prm.rho_d) / 2 * x + prm.rho / 2 * op_cldr(g, x)
    if prm.ablation == 'DGLR':
        return hx + prm.rho / 2 * op_cldr(g, x) + prm.rho_u / 2 * x
    return hx + (prm.rho_u + prm.rho_d) / 2 * x          # 'DGTV' and 'UT'


def lhs_zu(g, prm, z):
    """``LHS_zu`` (ADMM.py:389-390)."""
    return prm.mu_u * op_lu(g, z) + prm.rho_u / 2 * z


def lhs_zd(g, prm, z):
    """``LHS_zd`` (ADMM.py:392-394; the 'UT' branch at 395 is unreachable)."""
    return prm.mu_d2 * op_cldr(g, z) + prm.rho_d / 2 * z


def soft_phi(g, prm, x, gamma):
    """``phi_direct`` (ADMM.py:401-408)."""
    s = op_ldr(g, x) - gamma / prm.rho
    u = torch.abs(s) - prm.mu_d1 / prm.rho
    return torch.sign(s) * u * (u > 0)


def cg(apply_a, rhs, x0=None, max_iter=100, tol=1e-8, first_kwargs=None):
    """``CG_solver`` (ADMM.py:329-368).

    Per-window alpha/beta, batch-global stop test (quirk Q3).  ``first_kwargs`` go to the
    initial residual only (quirk Q4).  Returns ``(x, iters, alphas, betas)`` with
    ``iters = -1`` and plain lists when the tolerance was not reached; on convergence the
    lists are stacked to ``(iters, B)`` (the reference's ``torch.Tensor(list)`` only works
    for B = 1 — quirk Q2 — where it yields ``(iters,)``; we keep the trailing B)."""
    x = torch.zeros_like(rhs) if x0 is None else x0.clone()
    r = rhs - apply_a(x, **(first_kwargs or {}))
    p = r.clone()
    rr = (r * r).sum((1, 2, 3))
    alphas, betas = [], []
    for k in range(max_iter):
        ap = apply_a(p)
        a = rr / (p * ap).sum((1, 2, 3))
        alphas.append(a)
        x = x + a[:, None, None, None] * p
        r = r - a[:, None, None, None] * ap
        rr_new = (r * r).sum((1, 2, 3))
        b = rr_new / rr
        betas.append(b)
        rr = rr_new
        if torch.sqrt(rr).max() < tol:
            return x, k + 1, torch.stack(alphas), torch.stack(betas)
        p = r + b[:, None, None, None] * p
    return x, -1, alphas, betas


def first_guess(y, t_in, T):
    """``initial_guess`` (ADMM.py:766-781): per (window, node) least-squares line through the
    t_in observations, extrapolated to rows t_in..T-1.  The time axis is float32 whatever
    the dtype of y."""
    t = torch.arange(0, t_in, 1).to(torch.float)
    w = ((t[None, :, None, None] * y).mean(1) - t.mean() * y.mean(1)) / ((t ** 2).mean() - t.mean() ** 2)
    b = y.mean(1) - w * t.mean()
    t1 = torch.arange(t_in, T, 1).to(torch.float)
    return torch.cat((y, w[:, None, :, :] * t1[None, :, None, None] + b[:, None, :, :]), 1)


def first_interpolation(y, mask):
    """``initial_interpolation`` (ADMM.py:783-811)."""
    B, T, N, C = y.size()
    t = torch.arange(0, T, 1).to(torch.float).unsqueeze(0).unsqueeze(2).unsqueeze(3).repeat(B, 1, N, C)
    cnt = mask.sum(1)
    tm = (t * mask).sum(1) / cnt
    ym = (y * mask).sum(1) / cnt
    tym = (t * y * mask).sum(1) / cnt
    t2m = (t ** 2 * mask).sum(1) / cnt
    w = (tym - tm * ym) / (t2m - tm ** 2)
    b = ym - w * tm
    return (w * t + b) * (1 - mask) + y


@dataclass
class OracleTrace:
    """Everything ``combined_loop`` computes, including the iterates it keeps local."""
    x: torch.Tensor = None
    zu: torch.Tensor = None
    zd: torch.Tensor = None
    phi: torch.Tensor = None
    gamma: torch.Tensor = None
    gamma_u: torch.Tensor = None
    gamma_d: torch.Tensor = None
    cg_iter_x: list = field(default_factory=list)
    cg_iter_zu: list = field(default_factory=list)
    cg_iter_zd: list = field(default_factory=list)
    alpha_x: list = field(default_factory=list)
    beta_x: list = field(default_factory=list)
    alpha_zu: list = field(default_factory=list)
    beta_zu: list = field(default_factory=list)
    alpha_zd: list = field(default_factory=list)
    beta_zd: list = field(default_factory=list)
    p_res: list = field(default_factory=list)
    d_res: list = field(default_factory=list)
    x_shift: list = field(default_factory=list)
    delta_x_per_step: list = field(default_factory=list)
    glr: list = field(default_factory=list)
    dglr: list = field(default_factory=list)
    dgtv: list = field(default_factory=list)
    recover: list = field(default_factory=list)
    dx_mean: list = field(default_factory=list)      # (x - x_old).mean(0), the tensor behind delta_x_per_step
    outer_iters: int = 0


def admm_combined(g: OracleGraph, prm: OracleParams, y, mask=None, max_admm_iter=150, max_cg_iter=100,
                  cg_tol=1e-8, admm_tol=1e-6) -> OracleTrace:
    """``combined_loop`` (ADMM.py:511-648), ablation 'None' / 'DGTV' / 'DGLR' / 'UT'."""
    tr = OracleTrace()
    x = first_guess(y, prm.t_in, prm.T) if mask is None else first_interpolation(y, mask)
    assert not torch.isnan(g.d_w).any(), 'Directed graph weights d_ew has NaN value'
    assert not torch.isnan(g.u_w).any(), 'Undirected graph weights u_ew has NaN value'
    gu, gd = torch.ones_like(x) * 0.1, torch.ones_like(x) * 0.1
    with_phi = prm.ablation in ('None', 'DGLR')
    with_zd = prm.ablation != 'DGLR'
    gamma = phi = None
    if with_phi:
        gamma = torch.ones_like(x) * 0.1
        phi = op_ldr(g, x)
        assert not torch.isnan(phi).any(), 'initial phi has NaN value'
    zu, zd = x.clone(), x.clone()

    def solve(a, rhs, x0, **first):
        return cg(a, rhs, x0, max_iter=max_cg_iter, tol=cg_tol, first_kwargs=first)

    for i in range(max_admm_iter):
        x_old, zu_old, zd_old = x, zu, zd
        hty = torch.zeros_like(x)
        hty[:, 0:y.size(1)] = y
        if prm.ablation == 'None':                        # ADMM.py:559
            rhs = op_ldr_t(g, gamma + prm.rho * phi) / 2 + (prm.rho_u * zu + prm.rho_d * zd) / 2 - (gu + gd) / 2 + hty
        elif prm.ablation == 'DGLR':                      # ADMM.py:564
            rhs = op_ldr_t(g, gamma + prm.rho * phi) / 2 + prm.rho_u * zu / 2 - gu / 2 + hty
        else:                                             # ADMM.py:557, 562
            rhs = (prm.rho_u * zu + prm.rho_d * zd) / 2 - (gu + gd) / 2 + hty
        assert not torch.isnan(rhs).any(), f'RHS_x has NaN value in ADMM loop {i}'
        x, it, al, be = solve(lambda v, mask=None: lhs_x(g, prm, v, mask), rhs, x_old, mask=mask)
        tr.cg_iter_x.append(it); tr.alpha_x.append(al); tr.beta_x.append(be)
        assert not torch.isnan(x).any(), f'RHS_x has NaN value in loop {i}'
        assert not torch.isinf(x).any(), f'x has inf value in loop {i}'
        zu, it, al, be = solve(lambda v: lhs_zu(g, prm, v), gu / 2 + prm.rho_u / 2 * x, zu_old)
        tr.cg_iter_zu.append(it); tr.alpha_zu.append(al); tr.beta_zu.append(be)
        assert not torch.isnan(zu).any(), f'zu has NaN value in loop {i}'
        if with_zd:
            zd, it, al, be = solve(lambda v: lhs_zd(g, prm, v), gd / 2 + prm.rho_d / 2 * x, zd_old)
            tr.cg_iter_zd.append(it); tr.alpha_zd.append(al); tr.beta_zd.append(be)
            assert not torch.isnan(zd).any(), f'zd has NaN value in loop {i}'
        gu = gu + prm.rho_u * (x - zu)
        if with_zd:
            gd = gd + prm.rho_d * (x - zd)
        if with_phi:
            phi_old = phi.clone()
            phi = soft_phi(g, prm, x, gamma)
            assert not torch.isnan(phi).any(), f"phi has NaN value in loop {i}"
            gamma = gamma + prm.rho * (phi - op_ldr(g, x))
            assert not torch.isnan(gamma).any(), 'gamma has NaN'
        # diagnostics, ADMM.py:609-643
        pri, dual = [], []
        tr.x_shift.append(torch.norm(x - x_old).item())
        tr.dx_mean.append((x - x_old).mean(0))
        tr.delta_x_per_step.append((x - x_old).mean(0).norm(dim=[1, 2]))
        pri.append(torch.norm(x - zu).item())
        dual.append(torch.norm(zu - zu_old).item())
        tr.glr.append(glr(g, x))
        hx = x * mask if mask is not None else x[:, :prm.t_in]
        assert hx.size() == y.size(), f'Hx size {hx.size()}, y size {y.size()} not equal'
        tr.recover.append(torch.norm(hx - y).item())
        if with_phi:
            pri.append(torch.norm(phi - op_ldr(g, x)).item())
            dual.append(torch.norm(phi - phi_old).item())
            tr.dgtv.append(dgtv(g, x))
        if with_zd:
            pri.append(torch.norm(x - zd).item())
            dual.append(torch.norm(zd - zd_old).item())
            tr.dglr.append(dglr(g, x))
        tr.p_res.append(pri)
        tr.d_res.append(dual)
        tr.outer_iters = i + 1
        if max(pri) < admm_tol and max(dual) < admm_tol:
            break
    tr.x, tr.zu, tr.zd, tr.phi, tr.gamma, tr.gamma_u, tr.gamma_d = x, zu, zd, phi, gamma, gu, gd
    return tr


def two_loops(g: OracleGraph, prm: OracleParams, y, mask=None, max_admm_iter=150, max_inner_iter=100, max_cg_iter=100,
              cg_tol=1e-8) -> OracleTrace:
    """``two_loops`` (ADMM.py:410-508): an outer loop over the phi / gamma update around an inner loop of x, z_u, z_d
    solves whose duals gamma_u, gamma_d and splits z_u, z_d are RESET at every outer iteration (ADMM.py:446-447).
    The reference has no stop test and no residual lists here and returns nothing (the method ends without a
    ``return``); the trace keeps what its locals hold at the end and the CG lists it appends."""
    tr = OracleTrace()
    x = first_guess(y, prm.t_in, prm.T) if mask is None else first_interpolation(y, mask)
    with_phi = prm.ablation in ('None', 'DGLR')
    with_zd = prm.ablation != 'DGLR'
    gamma = phi = None
    if with_phi:
        gamma = torch.ones_like(x) * 0.1
        phi = op_ldr(g, x)
    gu = gd = zu = zd = None
    for _ in range(max_admm_iter):
        gu, gd = torch.ones_like(x) * 0.1, torch.ones_like(x) * 0.1
        zu, zd = x.clone(), x.clone()
        for _ in range(max_inner_iter):
            x_old, zu_old, zd_old = x, zu, zd
            hty = torch.zeros_like(x)
            hty[:, 0:y.size(1)] = y
            if prm.ablation == 'None':
                rhs = op_ldr_t(g, gamma + prm.rho * phi) / 2 + (prm.rho_u * zu + prm.rho_d * zd) / 2 - (gu + gd) / 2 + hty
            elif prm.ablation == 'DGLR':
                rhs = op_ldr_t(g, gamma + prm.rho * phi) / 2 + prm.rho_u * zu / 2 - gu / 2 + hty
            else:
                rhs = (prm.rho_u * zu + prm.rho_d * zd) / 2 - (gu + gd) / 2 + hty
            assert not torch.isnan(rhs).any(), 'RHS_x has NaN value in ADMM loop'                    # ADMM.py:466
            x, it, al, be = cg(lambda v, mask=None: lhs_x(g, prm, v, mask), rhs, x_old, max_iter=max_cg_iter, tol=cg_tol,
                               first_kwargs={"mask": mask})
            tr.cg_iter_x.append(it); tr.alpha_x.append(al); tr.beta_x.append(be)
            assert not torch.isnan(x).any(), 'RHS_x has NaN value in loop'                           # ADMM.py:473 (its wording)
            assert not torch.isinf(x).any(), 'x has inf value in loop'                               # ADMM.py:474
            zu, it, al, be = cg(lambda v: lhs_zu(g, prm, v), gu / 2 + prm.rho_u / 2 * x, zu_old, max_iter=max_cg_iter, tol=cg_tol)
            tr.cg_iter_zu.append(it); tr.alpha_zu.append(al); tr.beta_zu.append(be)
            assert not torch.isnan(zu).any(), 'zu has NaN value in loop'                             # ADMM.py:482
            if with_zd:
                zd, it, al, be = cg(lambda v: lhs_zd(g, prm, v), gd / 2 + prm.rho_d / 2 * x, zd_old, max_iter=max_cg_iter,
                                    tol=cg_tol)
                tr.cg_iter_zd.append(it); tr.alpha_zd.append(al); tr.beta_zd.append(be)
                assert not torch.isnan(zd).any(), 'zd has NaN value in loop'                         # ADMM.py:490
            gu = gu + prm.rho_u * (x - zu)
            if with_zd:
                gd = gd + prm.rho_d * (x - zd)
        if with_phi:
            phi = soft_phi(g, prm, x, gamma)
            gamma = gamma + prm.rho * (phi - op_ldr(g, x))
    tr.x, tr.zu, tr.zd, tr.phi, tr.gamma, tr.gamma_u, tr.gamma_d = x, zu, zd, phi, gamma, gu, gd
    return tr
