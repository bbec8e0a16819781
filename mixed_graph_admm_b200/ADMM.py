"""Drop-in for the reference's ``ADMM.py``: same class, method names, arguments and side effects,
with the inner loop running in hand-written sm_100a CUDA behind the C ABI of ``include/mga.h``.

What stays on the host (Python): argument checking, graph construction (once per instance),
the bookkeeping lists the reference fills (``ADMM.py:66-92``).  What moved: every operator
application, the three CG solves, the prox / dual updates and the diagnostics of
``combined_loop`` (``ADMM.py:511-648``).  There is no CPU path: without ``libmga.so`` and a
CUDA device every compute method raises.

Inputs may live on the CPU (as every caller of the reference passes them) or on the GPU;
results come back on the device of the input.
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import numpy as np
import torch

from . import _cabi
from .utils import *  # noqa: F401,F403  (the reference does `from utils import *`)
from .utils import (connect_list, directed_graph_from_distance, expand_time_dimension, get_data_difference,
                    k_nearest_neighbors, undirected_graph_from_distance)

__all__ = ["ADMM_algorithm", "initial_guess", "initial_interpolation"]


def _device_of(arg):
    if arg is not None:
        return torch.device(arg)
    if not torch.cuda.is_available():
        raise RuntimeError("mixed_graph_admm_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


@functools.lru_cache(maxsize=None)
def _regression_consts(t_in):
    """mean(t) and var(t) exactly as the reference forms them: float32 (ADMM.py:772-774)."""
    t = torch.arange(0, t_in, 1).to(torch.float)
    return float(t.mean()), float((t ** 2).mean() - t.mean() ** 2)


class _Plan:
    """Owns one ``mga_plan`` (device copies of the graph tables)."""

    def __init__(self, desc_kwargs, device):
        L = _cabi.lib()
        self.keep = []          # host tensors the descriptor points to
        d = _cabi.GraphDesc()
        for name, val in desc_kwargs.items():
            if isinstance(val, torch.Tensor):
                val = val.contiguous()
                self.keep.append(val)
                setattr(d, name, val.data_ptr())
            else:
                setattr(d, name, val)
        h = C.c_void_p()
        _cabi.check(L.mga_plan_create(C.byref(d), device.index or 0, C.byref(h)))
        self.handle = h
        self.device = device

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _cabi.lib().mga_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_RESULT_LISTS = ("alpha_x", "beta_x", "alpha_zu", "beta_zu", "alpha_zd", "beta_zd", "CG_iter_x", "CG_iter_zu", "CG_iter_zd",
                 "p_res_list", "d_res_list", "x_shift_list", "delta_x_per_step", "DGTV_list", "DGLR_list", "GLR_list",
                 "recover_list")


class _ResultList:
    """The result lists of the reference (ADMM.py:66-92) are plain python lists on the instance, and they are here too -
    but a call with ``print_info=False`` only records what it has to append (the diagnostics sums and the CG coefficient
    arrays, already on the host); the entries - a few hundred tensor views per call - are created when a list is first
    read.  Reading, appending to or replacing a list behaves exactly as with eager filling."""

    def __init__(self, name):
        self.slot = "_" + name

    def __get__(self, obj, cls):
        if obj is None:
            return self
        if obj._pending:
            obj._flush_pending()
        return obj.__dict__[self.slot]

    def __set__(self, obj, value):
        if obj.__dict__.get("_pending"):
            obj._flush_pending()
        obj.__dict__[self.slot] = value


class ADMM_algorithm():
    '''
    only with 1 head (reference ADMM.py:11-14)

    Extra keyword-only arguments (not in the reference): ``device`` (default: current CUDA
    device), ``mode`` ('auto' | 'streaming' | 'streaming_point' | 'resident'), ``verbose`` (print the sigma lines
    the reference prints from graph construction).
    '''

    for _name in _RESULT_LISTS:
        locals()[_name] = _ResultList(_name)
    del _name

    def __init__(self, graph_info, ADMM_info, use_kNN=False, k=4, u_sigma=None, d_sigma=None, expand_time_dim=True,
                 ablation='None', t_in=12, T=24, use_line_graph=False, skip_connection=1, *, device=None,
                 mode='auto', verbose=False):
        self._pending = []              # calls whose result-list entries have not been materialised yet (_ResultList)
        self.t_in = t_in
        self.T = T
        self.use_line_graph = use_line_graph
        self.skip_connection = skip_connection
        self.n_nodes = graph_info['n_nodes']
        self.u_edges = graph_info['u_edges']
        self.u_dists = graph_info['u_dist']
        self.use_kNN = use_kNN
        if use_kNN:
            self.connect_list, self.dist_list = k_nearest_neighbors(self.n_nodes, self.u_edges, self.u_dists, k)
            self.connect_list = self.connect_list.to(torch.int64)
        else:
            self.connect_list, self.dist_list = connect_list(self.n_nodes, self.u_edges, self.u_dists)

        self.ablation = ablation
        assert ablation in ['None', 'DGTV', 'DGLR', 'UT'], "ablation should be in ['None', 'DGTV', 'DGLR', 'UT']"
        self.u_ew = undirected_graph_from_distance(self.connect_list, self.dist_list, u_sigma=u_sigma,
                                                   regularized=True, verbose=verbose)
        if expand_time_dim:
            self.u_ew = expand_time_dimension(self.u_ew, T)
        if not self.use_line_graph:
            self.d_ew = directed_graph_from_distance(self.connect_list, self.dist_list, d_sigma=d_sigma,
                                                     regularized=True, verbose=verbose)
            if expand_time_dim:
                self.d_ew = expand_time_dimension(self.d_ew, T - 1)
        else:
            # banded temporal weights (T, skip, N) and source-time table (T, skip), ADMM.py:41-52
            w = torch.ones((self.n_nodes, self.T, self.skip_connection))
            w.tril_(diagonal=-1)
            w[:, 0, 0].fill_(1)
            w = w / w.sum(-1, keepdim=True)
            w[:, 0, 0].fill_(0)
            self.d_ew = w.permute(1, 2, 0)
            self.time_list = torch.arange(0, self.T).unsqueeze(1) - torch.arange(1, self.skip_connection + 1)

        self.rho = ADMM_info['rho']
        self.rho_u = ADMM_info['rho_u']
        self.rho_d = ADMM_info['rho_d']
        self.mu_u = ADMM_info['mu_u']
        self.mu_d1 = ADMM_info['mu_d1']
        self.mu_d2 = ADMM_info['mu_d2']

        self.max_CG_iter = 100
        self.max_inner_iter = 100
        self.CG_tol = 1e-8
        self.ADMM_tol = 1e-6
        self.max_ADMM_iter = 150

        self._reset_lists(all_lists=True)

        # ---- not in the reference
        self._device_arg = device       # resolved on first use: construction itself needs no GPU
        self._device = None
        self.mode = mode
        self.diag_reduce = None         # multi-GPU hook: (diag, dx_sum, B) -> global (diag, dx_sum, B)
        self.strict_quirks = False      # True: reproduce the B>1 converged-return ValueError (quirk Q2)
        self.keep_iterates = False      # True: combined_loop leaves z_u, z_d, phi, duals in last_iterates
        self.keep_cg_coefficients = True    # False: the CPU-input path skips the alpha_* / beta_* read-back (lists of [])
        self.last_iterates = None
        self.last_mode = None
        self._plan_key = None
        self._plan_obj = None
        self._plan_cache = {}

    @property
    def device(self):
        if self._device is None:
            self._device = _device_of(self._device_arg)
        return self._device

    @device.setter
    def device(self, value):
        self._device_arg, self._device = value, None

    # ------------------------------------------------------------------ bookkeeping
    def _flush_pending(self):
        pending, self._pending = self._pending, []
        for args in pending:
            self._fill_lists(*args)

    def _reset_lists(self, all_lists):
        if all_lists:
            self._pending = []          # every list is replaced: what was pending for them is void
        elif self._pending:
            self._flush_pending()       # recover_list survives init_iterations (ADMM.py:100-132)
        self.alpha_x = []
        self.beta_x = []
        self.alpha_zu = []
        self.beta_zu = []
        self.alpha_zd = []
        self.beta_zd = []
        self.CG_iter_x = []
        self.CG_iter_zu = []
        self.CG_iter_zd = []
        self.p_res_list = []
        self.d_res_list = []
        self.x_shift_list = []
        self.delta_x_per_step = []
        self.DGTV_list = []
        self.DGLR_list = []
        self.GLR_list = []
        if all_lists:
            self.recover_list = []      # init_iterations does not clear it (ADMM.py:100-132)
        self.res_name = ['zu']
        if self.ablation in ['None', 'DGLR']:
            self.res_name.append('phi')
        if self.ablation != 'DGLR':
            self.res_name.append('zd')

    def init_iterations(self, ablation, use_line_graph=False):
        """ADMM.py:100-132, including quirk Q8 (d_ew reset to default sigma, not time-expanded)."""
        if use_line_graph:
            self.use_line_graph = True
            self.d_ew = torch.ones((self.n_nodes, 1))
        else:
            self.use_line_graph = False
            self.d_ew = directed_graph_from_distance(self.connect_list, self.dist_list, d_sigma=None,
                                                     regularized=True)
        self.ablation = ablation
        self._reset_lists(all_lists=False)

    # ------------------------------------------------------------------ plan / params
    def _plan(self, Cn=1):
        """The device plan for signals with ``Cn`` channels.  The operators act on every channel with the same
        weights (``u_ew.unsqueeze(-1)``, ADMM.py:147,171) while dot products and norms run over (T, N, C)
        (ADMM.py:347-356), so a (T, N, C) window on the N-node graph IS a (T, N*C) window on the graph with
        node (i, c) -> i*C + c linked to (neighbour_j(i), c): the tables are expanded on the host and the
        kernels see one channel."""
        cl, uw, dw = self.connect_list, self.u_ew, self.d_ew
        key = (cl.data_ptr(), cl._version, tuple(cl.shape), uw.data_ptr(), uw._version, tuple(uw.shape),
               dw.data_ptr(), dw._version, tuple(dw.shape), self.t_in, self.T, self.use_line_graph,
               self.skip_connection, self.use_kNN, str(self.device), int(Cn))
        if key == self._plan_key:
            return self._plan_obj
        if key in self._plan_cache:                   # callers that alternate channel counts keep both plans
            self._plan_key, self._plan_obj = key, self._plan_cache[key]
            return self._plan_obj
        N, T = self.n_nodes, self.T
        cl64 = cl.detach().to('cpu', torch.int64)
        uw32 = uw.detach().to('cpu', torch.float32)
        dw32 = dw.detach().to('cpu', torch.float32)
        if tuple(cl64.shape[:1]) != (N,):
            raise ValueError(f"connect_list has {cl64.shape[0]} rows, graph has {N} nodes")
        if Cn > 1 and uw32.dim() >= 2 and dw32.dim() >= 2:
            ch = torch.arange(Cn, dtype=torch.int64).view(1, Cn, 1)
            cl64 = torch.where(cl64.unsqueeze(1) >= 0, cl64.unsqueeze(1) * Cn + ch,
                               cl64.unsqueeze(1).expand(N, Cn, cl64.shape[1])).reshape(N * Cn, -1)
            uw32 = uw32.repeat_interleave(Cn, dim=-2)
            # graph weights (N,K) / (T-1,N,K): node axis is -2; banded line-graph weights (T,skip,N): node axis is -1
            dw32 = dw32.repeat_interleave(Cn, dim=-1 if (self.use_line_graph and self.skip_connection > 1) else -2)
            N = N * Cn
        nbr_u = cl64[:, 1:].contiguous()
        if uw32.dim() == 2:
            u_T = 1
        elif uw32.dim() == 3 and uw32.shape[0] == T:
            u_T = T
        else:
            raise ValueError(f"u_ew must be (N,k) or (T,N,k), got {tuple(uw32.shape)}")
        if tuple(uw32.shape[-2:]) != (N, nbr_u.shape[1]):
            raise ValueError(f"u_ew shape {tuple(uw32.shape)} does not match connect_list {tuple(cl.shape)}")
        desc = dict(n_nodes=N, T=T, t_in=self.t_in, ku=nbr_u.shape[1], nbr_u=nbr_u, u_w=uw32, u_w_T=u_T,
                    ldrt_mode=_cabi.LDRT_SCATTER if self.use_kNN else _cabi.LDRT_GATHER)
        if not self.use_line_graph:
            if dw32.dim() == 2:
                d_T = 1
            elif dw32.dim() == 3 and dw32.shape[0] == T - 1:
                d_T = T - 1
            else:
                raise ValueError(f"d_ew must be (N,K) or (T-1,N,K), got {tuple(dw32.shape)}")
            if tuple(dw32.shape[-2:]) != tuple(cl64.shape):
                raise ValueError(f"d_ew shape {tuple(dw32.shape)} does not match connect_list {tuple(cl.shape)}")
            desc.update(kd=cl64.shape[1], nbr_d=cl64.contiguous(), d_w=dw32, d_w_T=d_T,
                        temporal=_cabi.TEMPORAL_GRAPH)
        elif self.skip_connection == 1:
            desc.update(kd=1, nbr_d=None, d_w=None, d_w_T=1, temporal=_cabi.TEMPORAL_LINE)
        else:
            if tuple(dw32.shape) != (T, self.skip_connection, N):
                raise ValueError("line graph with skip_connection > 1 needs d_ew of shape (T, skip, N)")
            desc.update(kd=self.skip_connection, nbr_d=None, d_w=dw32, d_w_T=1, temporal=_cabi.TEMPORAL_BAND)
        desc = {k: (0 if v is None else v) for k, v in desc.items()}
        self._plan_obj = _Plan(desc, self.device)
        # the key holds addresses: pin the tensors it was taken from so that a reassigned attribute (a sigma sweep
        # that builds a fresh d_ew) can never be given a recycled address that matches a cached entry
        self._plan_obj.key_tensors = (cl, uw, dw)
        self._plan_key = key
        if len(self._plan_cache) >= 4:
            self._plan_cache.clear()
        self._plan_cache[key] = self._plan_obj
        return self._plan_obj

    def _params(self):
        p = _cabi.Params()
        p.rho, p.rho_u, p.rho_d = float(self.rho), float(self.rho_u), float(self.rho_d)
        p.mu_u, p.mu_d1, p.mu_d2 = float(self.mu_u), float(self.mu_d1), float(self.mu_d2)
        p.ablation = _cabi.ABLATION[self.ablation]
        return p

    def _in(self, x):
        """(B,T,N,C) signal -> contiguous tensor on the plan's device."""
        if x.dim() != 4:
            raise ValueError(f"signals are (B, T, N, C); got {tuple(x.shape)}")
        if x.size(-1) < 1:
            raise ValueError("signals need at least one channel")
        if x.size(2) != self.n_nodes:
            raise ValueError(f"signal has {x.size(2)} nodes, graph has {self.n_nodes}")
        _cabi.dtype_id(x.dtype)
        return x.detach().to(self.device).contiguous()

    def _apply(self, op, x, mask=None):
        if x.size(1) != self.T:
            raise ValueError(f"operator input must have T={self.T} time steps, got {x.size(1)}")
        xd = self._in(x)
        md = self._in(mask.to(x.dtype)) if mask is not None else None
        y = torch.empty_like(xd)
        plan, prm = self._plan(xd.size(-1)), self._params()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().mga_apply(plan.handle, _cabi.OP[op], C.byref(prm), _cabi.ptr(xd), _cabi.ptr(y),
                                              _cabi.ptr(md), xd.size(0), _cabi.dtype_id(xd.dtype),
                                              _cabi.stream_ptr(self.device)))
        return y.to(x.device)

    # ------------------------------------------------------------------ operators (ADMM.py:138-228)
    def apply_op_Lu(self, x):
        '''signal shape: (B, T, N, n_channels)'''
        return self._apply("Lu", x)

    def apply_op_Ldr(self, x):
        return self._apply("Ldr", x)

    def apply_op_Ldr_T(self, x: torch.Tensor):
        return self._apply("Ldr_T", x)

    def apply_op_cLdr(self, x):
        return self._apply("cLdr", x)

    def apply_op_Ln(self, x):
        raise NotImplementedError("apply_op_Ln is unreachable in the reference (LHS_zd tests != 'DGLR' before "
                                  "== 'UT', ADMM.py:393-396) and is outside the hot path")

    def DGLR(self, x):
        '''x in (B, T, N, C); mean over the batch of ||L_d x||^2 (ADMM.py:230-235)'''
        return (self.apply_op_Ldr(x) ** 2).sum((1, 2, 3)).mean()

    def DGTV(self, x):
        '''mean over the batch of ||L_d x||_1 (ADMM.py:238-243)'''
        return self.apply_op_Ldr(x).norm(dim=[1, 2, 3], p=1).mean()

    def GLR(self, x):
        return (x * self.apply_op_Lu(x)).sum((1, 2, 3)).mean()

    # ------------------------------------------------------------------ systems (ADMM.py:371-408)
    def LHS_x(self, x, mask=None):
        return self._apply("LHS_x", x, mask)

    def LHS_zu(self, zu):
        return self._apply("LHS_zu", zu)

    def LHS_zd(self, zd):
        if self.ablation != 'DGLR':
            return self._apply("LHS_zd", zd)
        print('Error: LHS_zd')
        return None

    def phi_direct(self, x, gamma):
        '''phi = soft_(mu_d1 / rho) (L^d_r x - gamma / rho)   (ADMM.py:401-408)'''
        xd, gd = self._in(x), self._in(gamma)
        out = torch.empty_like(xd)
        plan, prm = self._plan(xd.size(-1)), self._params()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().mga_phi_direct(plan.handle, C.byref(prm), _cabi.ptr(xd), _cabi.ptr(gd),
                                                   _cabi.ptr(out), xd.size(0), _cabi.dtype_id(xd.dtype),
                                                   _cabi.stream_ptr(self.device)))
        return out.to(x.device)

    # ------------------------------------------------------------------ CG (ADMM.py:329-368)
    def _system_of(self, fn):
        owner = getattr(fn, "__self__", None)
        func = getattr(fn, "__func__", None)
        if owner is self:
            for name in ("x", "zu", "zd"):
                if func is getattr(type(self), "LHS_" + name):
                    return name
        return None

    def CG_solver(self, LHS_func, RHS, x0=None, **kwargs):
        '''
        Solving linear systems LHS_func(x) = RHS, B samples at the same time.
        ``LHS_func`` must be one of this instance's ``LHS_x`` / ``LHS_zu`` / ``LHS_zd`` for the fused
        CUDA path; any other callable runs the same recurrence with torch ops on the device.
        Returns ``(x, iters, alphas, betas)`` like the reference: on convergence ``iters = k+1`` and
        tensors, otherwise ``-1`` and python lists of ``(B,)`` tensors.
        '''
        system = self._system_of(LHS_func)
        if system is None:
            return self._cg_foreign(LHS_func, RHS, x0, **kwargs)
        if system == "zd" and self.ablation == 'DGLR':
            raise TypeError("LHS_zd returns None under ablation 'DGLR' (ADMM.py:397-399)")
        rhs = self._in(RHS)
        x = torch.zeros_like(rhs) if x0 is None else self._in(x0).clone()
        mask = kwargs.get("mask", None)
        md = self._in(mask.to(RHS.dtype)) if mask is not None else None
        B = rhs.size(0)
        n_it = int(self.max_CG_iter)
        alpha = torch.empty((max(n_it, 1), B), dtype=rhs.dtype, device=self.device)
        beta = torch.empty_like(alpha)
        iters = C.c_int32(-1)
        plan, prm = self._plan(rhs.size(-1)), self._params()
        with torch.cuda.device(self.device):
            # mode 'streaming' keeps the CG vectors in HBM; otherwise a fixed-iteration solve of a window
            # that fits one CTA runs in a single launch (include/mga.h: mga_plan_set_cg_mode)
            _cabi.check(_cabi.lib().mga_plan_set_cg_mode(
                plan.handle, _cabi.MODE[self.mode] if self.mode.startswith("streaming") else _cabi.MODE["auto"]))
            _cabi.check(_cabi.lib().mga_cg_solve(plan.handle, _cabi.SYS[system], C.byref(prm), _cabi.ptr(rhs),
                                                 _cabi.ptr(x), _cabi.ptr(md), B, _cabi.dtype_id(rhs.dtype), n_it,
                                                 float(self.CG_tol), C.byref(iters), _cabi.ptr(alpha),
                                                 _cabi.ptr(beta), _cabi.stream_ptr(self.device)))
        al, be = self._coef_lists(alpha, beta, iters.value, n_it, RHS.device)
        return x.to(RHS.device), iters.value, al, be

    def _coef_lists(self, alpha, beta, iters, n_it, device):
        """Types of quirk Q11: converged -> Tensor, else python list of (B,) tensors."""
        B = alpha.size(1)
        if iters > 0:
            if B > 1 and self.strict_quirks:
                raise ValueError("only one element tensors can be converted to Python scalars")   # ADMM.py:362
            a, b = alpha[:iters].to(device), beta[:iters].to(device)
            if B == 1:
                return a.reshape(-1).to(torch.float32), b.reshape(-1).to(torch.float32)   # torch.Tensor(list)
            return a, b
        a, b = alpha[:n_it].to(device), beta[:n_it].to(device)
        return list(a.unbind(0)), list(b.unbind(0))

    def _cg_foreign(self, LHS_func, RHS, x0=None, **kwargs):
        """The reference recurrence for a caller-supplied operator (not the fused path)."""
        alpha_list, beta_list = [], []
        x = torch.zeros_like(RHS) if x0 is None else x0.clone()
        r = RHS - LHS_func(x, **kwargs)
        p = r.clone()
        rr = (r * r).sum((1, 2, 3))
        for k in range(self.max_CG_iter):
            Ap = LHS_func(p)
            alpha = rr / (p * Ap).sum((1, 2, 3))
            alpha_list.append(alpha)
            x = x + alpha[:, None, None, None] * p
            r = r - alpha[:, None, None, None] * Ap
            rr_new = (r * r).sum((1, 2, 3))
            beta = rr_new / rr
            beta_list.append(beta)
            rr = rr_new
            if torch.sqrt(rr).max() < self.CG_tol:
                if rr.numel() > 1 and self.strict_quirks:
                    raise ValueError("only one element tensors can be converted to Python scalars")
                return x, k + 1, torch.stack(alpha_list).squeeze(-1), torch.stack(beta_list).squeeze(-1)
            p = r + beta[:, None, None, None] * p
        return x, -1, alpha_list, beta_list

    # ------------------------------------------------------------------ the loop (ADMM.py:511-648)
    def two_loops(self, y, mask=None, differential=False):
        '''
        two-loops algorithm (reference ADMM.py:410-508): an outer loop over the phi / gamma update around an inner loop
        (``max_inner_iter``) of x, z_u, z_d solves; gamma_u, gamma_d, z_u, z_d are reset at every outer iteration.
        Input:  y in (B, t_in, N, C)   [mask mode: y and mask in (B, T, N, C)]
        Like the reference it has no stop test, appends only the CG lists (``alpha_*``, ``beta_*``, ``CG_iter_*``) and
        RETURNS NOTHING (upstream the method ends without a ``return``); what its locals hold at the end is left in
        ``self.last_iterates`` (x, zu, zd, phi, gamma, gamma_u, gamma_d).  Every operator application, CG solve and prox
        goes through the C ABI (``mga_apply``, ``mga_cg_solve``, ``mga_phi_direct``) on device tensors.
        '''
        if differential:
            assert mask is None, 'differential mode does not support mask'   # (its first guess is discarded, ADMM.py:419-429)
        assert not torch.isnan(self.d_ew).any(), 'Directed graph weights d_ew has NaN value'
        assert not torch.isnan(self.u_ew).any(), 'Undirected graph weights u_ew has NaN value'
        out_device = y.device
        yd = self._in(y)
        md = self._in(mask.to(y.dtype)) if mask is not None else None
        x = initial_guess(yd, self.t_in, self.T) if mask is None else initial_interpolation(yd, md)
        with_phi = self.ablation in ['None', 'DGLR']
        with_zd = self.ablation != 'DGLR'
        gamma = phi = None
        if with_phi:
            gamma = torch.ones_like(x) * 0.1
            phi = self.apply_op_Ldr(x)
            assert not torch.isnan(phi).any(), 'initial phi has NaN value'
        gamma_u = gamma_d = zu = zd = None
        for _ in range(self.max_ADMM_iter):
            gamma_u, gamma_d = torch.ones_like(x) * 0.1, torch.ones_like(x) * 0.1
            zu, zd = x.clone(), x.clone()
            for i in range(self.max_inner_iter):
                x_old, zu_old, zd_old = x, zu, zd
                Hty = torch.zeros_like(x)
                Hty[:, 0:yd.size(1)] = yd
                if self.ablation == 'None':
                    RHS_x = self.apply_op_Ldr_T(gamma + self.rho * phi) / 2 + (self.rho_u * zu + self.rho_d * zd) / 2 \
                        - (gamma_u + gamma_d) / 2 + Hty
                elif self.ablation == 'DGLR':
                    RHS_x = self.apply_op_Ldr_T(gamma + self.rho * phi) / 2 + self.rho_u * zu / 2 - gamma_u / 2 + Hty
                else:
                    RHS_x = (self.rho_u * zu + self.rho_d * zd) / 2 - (gamma_u + gamma_d) / 2 + Hty
                assert not torch.isnan(RHS_x).any(), f'RHS_x has NaN value in ADMM loop {i}'
                x, CG_iter_x, alpha_x, beta_x = self.CG_solver(self.LHS_x, RHS_x, x_old, mask=md)
                self.alpha_x.append(alpha_x)
                self.beta_x.append(beta_x)
                self.CG_iter_x.append(CG_iter_x)
                assert not torch.isnan(x).any(), f'RHS_x has NaN value in loop {i}'
                assert not torch.isinf(x).any(), f'x has inf value in loop {i}'
                zu, CG_iter_zu, alpha_zu, beta_zu = self.CG_solver(self.LHS_zu, gamma_u / 2 + self.rho_u / 2 * x, zu_old)
                self.alpha_zu.append(alpha_zu)
                self.beta_zu.append(beta_zu)
                self.CG_iter_zu.append(CG_iter_zu)
                assert not torch.isnan(zu).any(), f'zu has NaN value in loop {i}'
                if with_zd:
                    zd, CG_iter_zd, alpha_zd, beta_zd = self.CG_solver(self.LHS_zd, gamma_d / 2 + self.rho_d / 2 * x, zd_old)
                    self.alpha_zd.append(alpha_zd)
                    self.beta_zd.append(beta_zd)
                    self.CG_iter_zd.append(CG_iter_zd)
                    assert not torch.isnan(zd).any(), f'zd has NaN value in loop {i}'
                gamma_u = gamma_u + self.rho_u * (x - zu)
                if with_zd:
                    gamma_d = gamma_d + self.rho_d * (x - zd)
            if with_phi:
                phi = self.phi_direct(x, gamma)
                assert not torch.isnan(phi).any(), "phi has NaN value"
                gamma = gamma + self.rho * (phi - self.apply_op_Ldr(x))
                assert not torch.isnan(gamma).any(), 'gamma has NaN'
        its = dict(x=x, zu=zu, zd=zd, phi=phi, gamma=gamma, gamma_u=gamma_u, gamma_d=gamma_d)
        self.last_iterates = {k: v.to(out_device) for k, v in its.items() if v is not None}
        self.last_mode = 'two_loops'
        return None

    def combined_loop(self, y, mask=None, differential=False, print_info=True):
        '''
        Input:  y in (B, t_in, N, C)   [mask mode: y and mask in (B, T, N, C)]
        Output: x in (B, T, N, C)
        '''
        if differential:
            assert mask is None, 'differential mode does not support mask'
            # the reference computes a differential first guess here and then discards it (ADMM.py:521-529)
        wkey = (self.d_ew.data_ptr(), self.d_ew._version, self.u_ew.data_ptr(), self.u_ew._version)
        if wkey != getattr(self, "_weights_checked", (None,))[0]:     # the reference re-checks every call (ADMM.py:517-518)
            assert not torch.isnan(self.d_ew).any(), 'Directed graph weights d_ew has NaN value'
            assert not torch.isnan(self.u_ew).any(), 'Undirected graph weights u_ew has NaN value'
            self._weights_checked = (wkey, self.d_ew, self.u_ew)    # the tensors stay alive: their addresses cannot be recycled
        out_device = y.device
        if y.dim() != 4:
            raise ValueError(f"signals are (B, T, N, C); got {tuple(y.shape)}")
        B, Cn = y.size(0), y.size(-1)
        T, N = self.T, self.n_nodes
        y_rows = y.size(1)
        if mask is None and y_rows != self.t_in:
            raise ValueError(f"y must have t_in={self.t_in} time steps, got {y_rows}")
        if mask is not None and (y_rows != T or tuple(mask.shape) != tuple(y.shape)):
            raise ValueError("mask mode needs y and mask of shape (B, T, N, C)")
        dt = _cabi.dtype_id(y.dtype)
        n_outer, n_cg = int(self.max_ADMM_iter), int(self.max_CG_iter)
        cg_tol, admm_tol = float(self.CG_tol), float(self.ADMM_tol)
        fixed = cg_tol <= 0 and admm_tol <= 0
        plan, prm = self._plan(Cn), self._params()
        L = _cabi.lib()
        t_mean, t_var = _regression_consts(self.t_in)
        want_iter = bool(self.keep_iterates)
        dev = self.device

        host_path = (y.device.type == 'cpu' and mask is None and fixed and not want_iter
                     and self.ablation == 'None')
        diag_h = dx_h = None
        cg_iters = np.full((max(n_outer, 1), 3), -1, dtype=np.int32)
        outer_done = C.c_int32(n_outer)
        alpha = beta = None
        with torch.cuda.device(dev):
            if host_path:
                # end-to-end call with host buffers: the copies overlap the solve (include/mga.h: mga_admm_solve_host);
                # the CG coefficients come back too, so the alpha_* / beta_* lists are what the reference fills
                yc = y.detach().contiguous()
                pin = yc.is_pinned()
                x = torch.empty((B, T, N, Cn), dtype=y.dtype, pin_memory=pin)
                # alpha | beta | diag | dx_sum in ONE host block: the library then brings them back in one copy
                want_coef = self.keep_cg_coefficients and n_cg > 0 and n_outer > 0
                coef_b = n_outer * 3 * n_cg * B * y.element_size() if want_coef else 0
                diag_n, dx_n = n_outer * _cabi.DIAG_COLS, n_outer * T * N * Cn
                blob = torch.empty((2 * coef_b + 8 * (diag_n + dx_n),), dtype=torch.uint8, pin_memory=pin)   # (2 * coef_b: a multiple of 8)
                if want_coef:
                    alpha = blob[:coef_b].view(y.dtype).view(n_outer, 3, n_cg, B)
                    beta = blob[coef_b:2 * coef_b].view(y.dtype).view(n_outer, 3, n_cg, B)
                dd = blob[2 * coef_b:].view(torch.float64)
                diag_t, dx_t = dd[:diag_n], dd[diag_n:]
                _cabi.check(L.mga_admm_solve_host(plan.handle, C.byref(prm), _cabi.ptr(yc), y_rows, _cabi.ptr(x), B,
                                                  dt, n_outer, n_cg, t_mean, t_var, 1, _cabi.ptr(diag_t), _cabi.ptr(dx_t),
                                                  _cabi.ptr(alpha), _cabi.ptr(beta), _cabi.MODE[self.mode], 0))
                diag_h = diag_t.numpy().reshape(n_outer, _cabi.DIAG_COLS)
                dx_h = dx_t.numpy().reshape(n_outer, T, N * Cn)
                self.last_mode = 'host'
                self.last_iterates = None
            else:
                yd = self._in(y)
                md = self._in(mask.to(y.dtype)) if mask is not None else None
                x = torch.empty((B, T, N, Cn), dtype=y.dtype, device=dev)
                outs = _cabi.AdmmOutputs()
                its = {}
                if want_iter:
                    for name in ("zu", "zd", "phi", "gamma", "gamma_u", "gamma_d"):
                        its[name] = torch.zeros_like(x)
                        setattr(outs, name, its[name].data_ptr())
                # (the library zeroes the diagnostics itself; alpha / beta are read up to the iteration counts only)
                no, dn, xn = max(n_outer, 1), max(n_outer, 1) * _cabi.DIAG_COLS, max(n_outer, 1) * T * N * Cn
                dd = torch.empty((dn + xn,), dtype=torch.float64, device=dev)
                diag_d, dx_d = dd[:dn].view(no, _cabi.DIAG_COLS), dd[dn:].view(no, T, N * Cn)
                alpha = torch.empty((2, no, 3, max(n_cg, 1), B), dtype=y.dtype, device=dev)
                alpha, beta = alpha[0], alpha[1]
                outs.diag, outs.dx_sum = diag_d.data_ptr(), dx_d.data_ptr()
                outs.alpha, outs.beta = alpha.data_ptr(), beta.data_ptr()
                outs.cg_iters = cg_iters.ctypes.data
                outs.outer_done = C.addressof(outer_done)
                _cabi.check(L.mga_admm_solve(plan.handle, C.byref(prm), _cabi.ptr(yd), y_rows, _cabi.ptr(md),
                                             _cabi.ptr(x), B, dt, n_outer, n_cg, cg_tol, admm_tol, t_mean, t_var, 1,
                                             C.byref(outs), _cabi.MODE[self.mode], _cabi.stream_ptr(dev)))
                dd_h = dd.cpu().numpy()                                   # (synchronises the stream)
                diag_h, dx_h = dd_h[:dn].reshape(no, _cabi.DIAG_COLS), dd_h[dn:].reshape(no, T, N * Cn)
                self.last_iterates = {k: v.to(out_device) for k, v in its.items()} if want_iter else None
                self.last_mode = 'device'
        if self.diag_reduce is not None:      # shards of one batch: sum the partial sums over the ranks
            diag_h, dx_h, B = self.diag_reduce(diag_h, dx_h, B)
            # (the per-window CG coefficients stay with their shard: alpha_* / beta_* hold (B_local,) tensors)
        n_done = int(outer_done.value)
        if n_done > 0 and diag_h[:n_done, _cabi.DIAG_NONFINITE].any():
            i = int(np.nonzero(diag_h[:n_done, _cabi.DIAG_NONFINITE])[0][0])
            raise AssertionError(f'x / z / phi / gamma has NaN or inf value in loop {i}')   # ADMM.py:575-606
        if self._pending:
            self._flush_pending()       # keep the order of the entries (and at most one call's arrays alive)
        args = (diag_h, dx_h, cg_iters, n_done, alpha, beta, B, y.dtype, out_device, print_info)
        if print_info:
            self._fill_lists(*args)     # the per-iteration lines are printed now
        else:
            self._pending.append(args)
        return x.to(out_device)

    def _fill_lists(self, diag, dx_sum, cg_iters, n_done, alpha, beta, B, dtype, device, print_info):
        """Append one entry per executed outer iteration to the reference's result lists
        (ADMM.py:572-643), with the element types of quirk Q11."""
        with_phi = self.ablation in ['None', 'DGLR']
        with_zd = self.ablation != 'DGLR'

        np_dt = np.float32 if dtype == torch.float32 else np.float64

        def rnd(v):      # value rounded to the signal dtype, as a python float
            return float(np_dt(v))

        # the tensor-valued entries (quirk Q11) of all iterations at once: 0-dim / (T,) views of four tensors
        nd = max(n_done, 1)
        means = torch.from_numpy((diag[:nd][:, [_cabi.DIAG_GLR, _cabi.DIAG_DGTV, _cabi.DIAG_DGLR]] / B).astype(np_dt)).to(device)
        mean_dx = dx_sum[:nd] / B
        dxs = torch.from_numpy(np.sqrt((mean_dx * mean_dx).sum(2)).astype(np_dt)).to(device)

        # fixed iteration counts (nothing converged): every entry of alpha_* / beta_* is a python list of n_cg (B,)
        # tensors (quirk Q11) - all rows of the (n_outer, 3, n_cg, B) arrays are unbound in ONE call each
        rows_a = rows_b = None
        if alpha is not None and n_done > 0 and bool((cg_iters[:n_done] < 0).all()):
            n_row = alpha.size(2)
            rows_a = alpha.to(device).reshape(-1, alpha.size(-1)).unbind(0)
            rows_b = beta.to(device).reshape(-1, beta.size(-1)).unbind(0)

        for i in range(n_done):
            d = diag[i]
            if d[_cabi.DIAG_NONFINITE] > 0:
                raise AssertionError(f'x / z / phi / gamma has NaN or inf value in loop {i}')   # ADMM.py:575-606
            its = [int(v) for v in cg_iters[i]]
            names = ["x", "zu"] + (["zd"] if with_zd else [])
            for s, name in enumerate(names):
                getattr(self, "CG_iter_" + name).append(its[s])
                if rows_a is not None:
                    r0 = (i * 3 + s) * n_row
                    a, b = list(rows_a[r0:r0 + n_row]), list(rows_b[r0:r0 + n_row])
                elif alpha is not None:
                    n_it = its[s] if its[s] > 0 else int(self.max_CG_iter)
                    a, b = self._coef_lists(alpha[i, s], beta[i, s], its[s], n_it, device)
                else:
                    a, b = [], []          # keep_cg_coefficients = False
                getattr(self, "alpha_" + name).append(a)
                getattr(self, "beta_" + name).append(b)
            pri, dual = [], []
            self.x_shift_list.append(rnd(math.sqrt(d[_cabi.DIAG_DX2])))
            self.delta_x_per_step.append(dxs[i])
            pri.append(rnd(math.sqrt(d[_cabi.DIAG_X_ZU2])))
            dual.append(rnd(math.sqrt(d[_cabi.DIAG_DZU2])))
            self.GLR_list.append(means[i, 0])
            self.recover_list.append(rnd(math.sqrt(d[_cabi.DIAG_RECOVER2])))
            if with_phi:
                pri.append(rnd(math.sqrt(d[_cabi.DIAG_PHI_LDX2])))
                dual.append(rnd(math.sqrt(d[_cabi.DIAG_DPHI2])))
                self.DGTV_list.append(means[i, 1])
            if with_zd:
                pri.append(rnd(math.sqrt(d[_cabi.DIAG_X_ZD2])))
                dual.append(rnd(math.sqrt(d[_cabi.DIAG_DZD2])))
                self.DGLR_list.append(means[i, 2])
            if print_info:
                zd_it = its[2] if with_zd else None
                print(f'ADMM iters {i}: x_CG_iters {its[0]}, zu_CG_iters {its[1]}, zd_CG_iters {zd_it}, '
                      f'pri_err = [{", ".join([f"{err:.4g}" for err in pri])}], '
                      f'dual_err = [{", ".join([f"{err:.4g}" for err in dual])}]')
            self.p_res_list.append(pri)
            self.d_res_list.append(dual)

    # plots (ADMM.py:650-761) read the lists above; matplotlib is not a dependency of the hot path
    def _no_plot(self, *a, **k):
        raise NotImplementedError("plotting is outside the hot path; the lists it reads (p_res_list, d_res_list, "
                                  "x_shift_list, delta_x_per_step, GLR/DGLR/DGTV_list, alpha_*/beta_*) are filled")

    plot_residual = plot_x_per_step = plot_CG_params = plot_regularization_terms = _no_plot


def initial_guess(y, t_in, T):
    '''
    y in (B, t_in, N, C) -> x in (B, T, N, C): per (window, node) least-squares line through the
    observations, extrapolated (ADMM.py:766-781).  Runs the CUDA kernel ``mga_initial_guess``.
    '''
    if y.size(1) != t_in:
        raise ValueError("y must have t_in time steps")
    dev = y.device if y.is_cuda else _device_of(None)
    B, Cn = y.size(0), y.size(3)
    N = y.size(2) * Cn          # the line is fitted per (window, node, channel): channels are extra nodes
    # a graph-free plan: initial_guess only needs the shape
    desc = dict(n_nodes=N, T=T, t_in=t_in, ku=0, nbr_u=0, u_w=0, u_w_T=1, kd=1, nbr_d=0, d_w=0, d_w_T=1,
                ldrt_mode=0, temporal=_cabi.TEMPORAL_LINE)
    plan = _Plan(desc, dev)
    yd = y.detach().to(dev).contiguous()
    x = torch.empty((B, T, y.size(2), Cn), dtype=y.dtype, device=dev)
    t_mean, t_var = _regression_consts(t_in)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mga_initial_guess(plan.handle, _cabi.ptr(yd), _cabi.ptr(x), B,
                                                  _cabi.dtype_id(y.dtype), t_mean, t_var, _cabi.stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()
    return x.to(y.device)


def initial_interpolation(y, mask):
    '''
    y, mask in (B, T, N, C), y = x * mask: per-node regression through the observed entries, used to
    fill the missing ones (ADMM.py:783-811).  Host-side torch ops on the device of ``y``; inside
    combined_loop the same step runs in the CUDA prologue kernel.  Unlike the reference (whose
    ``w * t`` broadcast only lines up for B == 1, ADMM.py:802) this works for any B.
    '''
    B, T, N, Cn = y.size()
    t = torch.arange(0, T, 1, device=y.device).to(torch.float).view(1, T, 1, 1).expand(B, T, N, Cn)
    n_data = mask.sum(1, keepdim=True)
    t_mean = (t * mask).sum(1, keepdim=True) / n_data
    y_mean = (y * mask).sum(1, keepdim=True) / n_data
    ty_mean = (t * y * mask).sum(1, keepdim=True) / n_data
    t2_mean = (t ** 2 * mask).sum(1, keepdim=True) / n_data
    w = (ty_mean - t_mean * y_mean) / (t2_mean - t_mean ** 2)
    b = y_mean - w * t_mean
    assert not torch.isnan(w).any(), 'Initial interpolation w has NaN value'
    assert not torch.isnan(b).any(), 'Initial interpolation b has NaN value'
    x = (w * t + b) * (1 - mask) + y
    assert not torch.isnan(x).any(), 'Initial interpolation x has NaN value'
    return x
