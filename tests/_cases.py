"""Shared helpers: load a golden fixture, rebuild the oracle / the CUDA solver from it."""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIXED_CASES = ["anchor5", "tiny_f32", "tiny_f64", "tiny_prox", "tiny_noexpand", "tiny_physical", "tiny_line1", "tiny_line2",
               "tiny_reinit", "tiny_abl_dgtv", "tiny_mask", "tiny_mask_f64", "tiny_c2", "tiny_c2_f64", "tiny_c3_line2", "tiny_c3_physical", "tiny_c2_mask",
               "pems08_f32", "pems04_f32"]
TOL_CASES = ["tiny_tol", "tiny_tol_f64", "tiny_c2_tol", "pems08_tol", "pems04_tol", "pems04_t24_tol_f64", "tiny_mask_tol_f64"]
ALL_CASES = FIXED_CASES + TOL_CASES
ITERATES = ("x", "zu", "zd", "phi", "gamma", "gamma_u", "gamma_d")


class Golden:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        self.ctor = self.meta["ctor"]
        self.limits = self.meta["limits"]
        self.admm_info = self.meta["admm_info"]
        self.init = self.meta["init"]
        self.dtype = getattr(torch, self.meta["dtype"])
        self.graph_info = {"n_nodes": self.meta["n_nodes"], "u_edges": torch.from_numpy(z["u_edges"]),
                           "u_dist": torch.from_numpy(z["u_dist"])}

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def has(self, key):
        return key in self.z.files

    @property
    def y(self):
        return self.t("y")

    @property
    def mask(self):
        return self.t("mask") if self.has("mask") else None

    def limit(self, key, default):
        return self.limits.get(key, default)


def oracle_from_golden(g: Golden):
    """OracleGraph / OracleParams from the tables the reference built (stored in the fixture)."""
    from oracle import admm_oracle as O
    line = bool(g.ctor.get("use_line_graph", False)) or (g.init is not None and g.init[1])
    og = O.OracleGraph(nbr=g.t("connect_list"), u_w=g.t("u_ew"), d_w=g.t("d_ew"),
                       use_knn=bool(g.ctor.get("use_kNN", False)), line_graph=line,
                       skip=int(g.ctor.get("skip_connection", 1)),
                       time_list=g.t("time_list") if g.has("time_list") else None)
    abl = g.init[0] if g.init is not None else g.ctor.get("ablation", "None")
    prm = O.OracleParams(**g.admm_info, t_in=g.ctor["t_in"], T=g.ctor["T"], ablation=abl)
    return og, prm


def run_oracle(g: Golden):
    from oracle import admm_oracle as O
    og, prm = oracle_from_golden(g)
    return O.admm_combined(og, prm, g.y, mask=g.mask, max_admm_iter=g.limit("max_ADMM_iter", 150),
                           max_cg_iter=g.limit("max_CG_iter", 100), cg_tol=g.limit("CG_tol", 1e-8),
                           admm_tol=g.limit("ADMM_tol", 1e-6))


def solver_from_golden(g: Golden, mode="auto", device=None):
    """The CUDA drop-in, constructed exactly like the reference was for this fixture."""
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    with contextlib.redirect_stdout(io.StringIO()):
        blk = ADMM_algorithm(g.graph_info, g.admm_info, **g.ctor, device=device, mode=mode)
        if g.init is not None:
            blk.init_iterations(*g.init)
    for k, v in g.limits.items():
        setattr(blk, k, v)
    return blk


def rel_err(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def max_rel(a, b, scale=None):
    """max |a-b| / max |b|: relative error in the sup norm (``scale`` overrides the denominator)."""
    a = a.double().flatten()
    b = b.double().flatten()
    den = b.abs().max().item() if scale is None else scale
    return ((a - b).abs().max() / max(den, 1e-300)).item()
