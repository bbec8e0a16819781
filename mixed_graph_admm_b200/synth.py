"""Seeded synthetic inputs for tests and benchmarks.

The PEMS datasets the reference was run on are not distributable, so every test and
benchmark here uses a synthetic road-like graph and random signal windows of the same
shape.  The dictionaries produced match what the reference's ``TrafficDataset`` hands to
``ADMM_algorithm`` (``utils.py:41-52``, ``utils.py:74-79``): bidirectional ``u_edges`` of
shape ``(2E, 2)`` int64 and ``u_dist`` of shape ``(2E,)`` float64.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def road_graph(n_nodes: int, edge_ratio: float = 1.1, seed: int = 0, isolate_pair: bool = False):
    """Planar-ish road graph: uniform points, Euclidean spanning tree + short extra links.

    The spanning tree of the 8-nearest-neighbour graph keeps the network (almost always)
    connected, like a road network; the shortest remaining nearest-neighbour links are then
    added until ``edge_ratio * N`` undirected edges exist (PEMS04: 340 / 307 = 1.11,
    PEMS08: 295 / 170 = 1.74).  Edge cost is the Euclidean length scaled into the PEMS range
    (about 3 .. 600) and rounded to 0.1.  With ``isolate_pair`` the last two nodes are cut
    off into their own 2-node component so kNN tables get ``-1`` padding.
    """
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree
    from scipy.spatial import cKDTree

    rng = np.random.RandomState(seed)
    pts = rng.rand(n_nodes, 2)
    n_main = n_nodes - 2 if isolate_pair else n_nodes
    kq = min(9, n_main)
    dist, idx = cKDTree(pts[:n_main]).query(pts[:n_main], k=kq)
    rows = np.repeat(np.arange(n_main), kq - 1)
    cols = idx[:, 1:].reshape(-1)
    vals = dist[:, 1:].reshape(-1)
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    cand = {}
    for a, b, d in zip(lo.tolist(), hi.tolist(), vals.tolist()):
        cand.setdefault((a, b), d)
    ck = sorted(cand.keys())
    cm = coo_matrix((np.array([cand[k] for k in ck]), ([k[0] for k in ck], [k[1] for k in ck])),
                    shape=(n_main, n_main))
    mst = minimum_spanning_tree(cm).tocoo()
    edges = {}
    for a, b, d in zip(mst.row.tolist(), mst.col.tolist(), mst.data.tolist()):
        edges[(min(a, b), max(a, b))] = d
    target = int(round(edge_ratio * n_nodes))
    for d, a, b in sorted((cand[k], k[0], k[1]) for k in ck):
        if len(edges) >= target:
            break
        edges.setdefault((a, b), d)
    if isolate_pair:
        a, b = n_nodes - 2, n_nodes - 1
        edges[(a, b)] = float(np.linalg.norm(pts[a] - pts[b]))
    keys = sorted(edges.keys())
    lens = np.array([edges[k] for k in keys])
    # scale so that the median link is ~60 units, clip into [3, 600], round to 0.1
    scale = 60.0 / max(np.median(lens), 1e-12)
    cost = np.round(np.clip(lens * scale, 3.0, 600.0), 1)
    frm = [k[0] for k in keys]
    to = [k[1] for k in keys]
    u_edges = torch.tensor([frm + to, to + frm], dtype=torch.int64).T.contiguous()
    u_dist = torch.tensor(np.concatenate([cost, cost]), dtype=torch.float64)
    return {"n_nodes": n_nodes, "n_edges": len(keys), "u_edges": u_edges, "u_dist": u_dist}


def admm_info(n_nodes: int):
    """Penalty / regulariser weights of the notebooks: rho_init = sqrt(N / 24)
    (``example-PEMS04.ipynb`` cell 8; SURVEY.md §8d fixes the 24 for every T)."""
    r = math.sqrt(n_nodes / 24.0)
    return {"rho": 2 * r, "rho_u": 3 * r, "rho_d": 2 * r, "mu_u": 1.0, "mu_d1": 2.0, "mu_d2": 1.0}


def signals(batch: int, t_in: int, n_nodes: int, seed: int = 0, dtype=torch.float32, smooth: bool = False):
    """Observation windows ``y`` of shape ``(B, t_in, N, 1)`` in [0, 1]."""
    g = torch.Generator().manual_seed(seed)
    if not smooth:
        return torch.rand(batch, t_in, n_nodes, 1, generator=g, dtype=dtype)
    t = torch.arange(t_in, dtype=torch.float64)[None, :, None, None]
    phase = torch.rand(batch, 1, 1, 1, generator=g, dtype=torch.float64) * 2 * math.pi
    offs = torch.rand(1, 1, n_nodes, 1, generator=g, dtype=torch.float64) * 0.4
    y = 0.3 + offs + 0.25 * torch.sin(2 * math.pi * t / 288.0 * 12 + phase)
    y = y + 0.05 * torch.rand(batch, t_in, n_nodes, 1, generator=g, dtype=torch.float64)
    return y.clamp_(0, 1).to(dtype)
