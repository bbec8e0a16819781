"""Host-side graph construction — the boundary input of the solver hot path.

Same function names, argument meaning and return layouts as the reference's ``utils.py``
(the step *before* the hot path), written from scratch:

* ``k_nearest_neighbors`` replaces the reference's N full networkx Dijkstra runs
  (``utils.py:183-204``) by a truncated Dijkstra that settles only ``k+1`` nodes per source
  and reproduces networkx's heap tie-breaking, so the tables are bit-identical
  (SURVEY.md §8c).  A native C++ version of the same search is used when the extension is
  built (``mga_knn_build``); this pure-Python one is the portable definition.
* the weight functions reproduce ``utils.py:206-258`` on the same torch ops so the float32
  weights are bit-identical.

No printing: the reference prints sigma / degree tables from these functions; pass
``verbose=True`` to get the same lines.
"""
from __future__ import annotations

import heapq
from itertools import count

import torch

__all__ = [
    "k_nearest_neighbors", "connect_list", "undirected_graph_from_distance",
    "directed_graph_from_distance", "expand_time_dimension", "line_graph",
    "get_data_difference",
]


def _adjacency(edges, dists):
    """Successor lists in first-insertion order, last weight wins — what
    ``nx.DiGraph.add_edge`` called edge by edge produces (``utils.py:190-192``)."""
    e = edges.tolist() if hasattr(edges, "tolist") else [list(x) for x in edges]
    d = dists.tolist() if hasattr(dists, "tolist") else list(dists)
    succ: dict[int, dict[int, float]] = {}
    for (a, b), w in zip(e, d):
        a = int(a)
        b = int(b)
        succ.setdefault(a, {})[b] = w
        succ.setdefault(b, {})
    return succ


def _settle_first(succ, source, n_settle):
    """First ``n_settle`` nodes a networkx Dijkstra from ``source`` would settle, in order.

    networkx keeps heap entries ``(dist, push_counter, node)`` and relaxes successors in
    adjacency order; ``heapq.nsmallest(k+1, dist.items(), key=dist)`` (``utils.py:200``) is
    stable, and the dict is already in non-decreasing settle order, so the kNN row is the
    first ``k+1`` settled nodes.  Distances accumulate in Python floats (float64).
    """
    if source not in succ:
        raise KeyError(f"Node {source} not found in graph")
    done: dict[int, float] = {}
    seen = {source: 0}
    c = count()
    heap = [(0, next(c), source)]
    while heap and len(done) < n_settle:
        d, _, v = heapq.heappop(heap)
        if v in done:
            continue
        done[v] = d
        for u, w in succ[v].items():
            if u in done:
                continue
            nd = d + w
            if u not in seen or nd < seen[u]:
                seen[u] = nd
                heapq.heappush(heap, (nd, next(c), u))
    return list(done.items())


def k_nearest_neighbors(n_nodes, edges: torch.Tensor, dists: torch.Tensor, k, verbose=False, *, native=None):
    """kNN by shortest-path distance (reference ``utils.py:183-204``).

    Returns ``(nearest_nodes (N, k+1) int32, nearest_dists (N, k+1) float32)`` with the node
    itself in column 0, ``-1`` / ``inf`` where fewer than ``k+1`` nodes are reachable.
    ``native`` (not in the reference): ``None`` uses the C++ builder of libmga when it is built,
    ``False`` forces the pure-Python search (bench.py's reference arm: nothing of the product
    is loaded there), ``True`` requires the C++ builder.  Both produce the same bits.
    """
    if native is None or native:
        tables = _native_knn(n_nodes, edges, dists, k)
        if tables is not None:
            return tables
        if native:
            raise RuntimeError("k_nearest_neighbors(native=True): libmga.so is not built")
    succ = _adjacency(edges, dists)
    if verbose:
        print(f'{n_nodes} nodes, {k} neighbors')
    nodes = -torch.ones((n_nodes, k + 1), dtype=torch.int)
    nd = torch.full((n_nodes, k + 1), float('inf'))
    for s in range(n_nodes):
        row = _settle_first(succ, s, k + 1)
        m = len(row)
        nodes[s, :m] = torch.tensor([a for a, _ in row])
        nd[s, :m] = torch.tensor([b for _, b in row])
    return nodes, nd


def _native_knn(n_nodes, edges, dists, k):
    """C++ truncated Dijkstra from the extension, if it is built; else None."""
    try:
        from . import _cabi
    except Exception:
        return None
    if not _cabi.available() or not hasattr(_cabi, "knn_build"):
        return None
    return _cabi.knn_build(n_nodes, edges, dists, k)


def connect_list(n_nodes, edges, dists, verbose=False):
    """Physical-adjacency table (reference ``utils.py:156-181``): ``(N, maxdeg+1)`` int64 with
    the node itself in column 0 and its out-neighbours stored in *reverse* edge order
    (the reference fills slot ``counts[i]`` and counts down), ``-1`` / ``inf`` padding."""
    e = torch.as_tensor(edges).to(torch.int64)
    src = e[:, 0]
    deg = torch.bincount(src, minlength=n_nodes)
    kmax = int(deg.max().item()) if e.numel() else 0
    if verbose:
        print(deg.to(torch.int))
        print('max degrees', kmax)
    table = -torch.ones((n_nodes, kmax + 1), dtype=torch.int64)
    dtab = torch.full((n_nodes, kmax + 1), float('inf'))
    left = deg.clone()
    dd = torch.as_tensor(dists)
    for i in range(e.shape[0]):
        a = int(src[i])
        slot = int(left[a])
        table[a, slot] = e[i, 1]
        dtab[a, slot] = dd[i]
        left[a] -= 1
    table[:, 0] = torch.arange(n_nodes)
    dtab[:, 0] = 0
    return table, dtab


def _default_sigma(table, dtab):
    live = (table != -1) & (dtab != 0)
    vals = dtab[live]
    lo, hi = vals.min().item(), vals.max().item()
    return max(hi / 50, lo * 50), lo, hi


def undirected_graph_from_distance(connect_list: torch.Tensor, dist_list: torch.Tensor, u_sigma=None,
                                   regularized=True, verbose=False):
    """Spatial weights ``(N, k)`` (reference ``utils.py:206-238``): ``exp(-d/sigma)`` on columns
    1.., zero where the neighbour is ``-1``, normalised by ``1/sqrt(deg_i * deg_j)`` where each
    degree is the sum of that row's *own* list only — the resulting Laplacian is not symmetric
    when kNN lists are not mutual (quirk Q5) and must stay that way."""
    n = connect_list.shape[0]
    sig_default, lo, hi = _default_sigma(connect_list, dist_list)
    if u_sigma == None:  # noqa: E711  (the reference accepts 0-dim tensors here)
        u_sigma = sig_default
    if verbose:
        print(f'Undirected graph: sigma = {u_sigma}, nearest_dist in ({lo:.4f}, {hi:.4f})')
    nb = connect_list[:, 1:]
    w = torch.exp(-dist_list[:, 1:] / u_sigma)
    w[nb == -1] = 0
    if regularized:
        deg = w.sum(1)
        pair = deg.unsqueeze(1) * deg[nb].reshape(n, -1)   # index -1 wraps, as in the reference
        scale = torch.where(pair > 0, 1 / torch.sqrt(pair), torch.zeros_like(pair))
        w = w * scale
    return w


def directed_graph_from_distance(connect_list: torch.Tensor, dist_list: torch.Tensor, d_sigma=None,
                                 regularized=True, verbose=False):
    """Temporal (directed) weights ``(N, k+1)`` incl. the self link (reference
    ``utils.py:240-258``): ``exp(-d/sigma)``, zero at ``-1``, rows normalised to sum 1."""
    sig_default, lo, hi = _default_sigma(connect_list, dist_list)
    if d_sigma == None:  # noqa: E711
        d_sigma = sig_default
    if verbose:
        print(f'Directed Graph: sigma = {d_sigma}, nearest_dist in ({lo:.4f}, {hi:.4f})')
    w = torch.exp(-dist_list / d_sigma)
    w[connect_list == -1] = 0
    if regularized:
        tot = w.sum(1)
        inv = torch.where(tot > 0, 1 / tot, torch.zeros_like(tot))
        w = w * inv.unsqueeze(1)
    return w


def expand_time_dimension(ew, T: int):
    """``(N, k) -> (T, N, k)`` by repetition (reference ``utils.py:294-295``)."""
    return ew.unsqueeze(0).repeat(T, 1, 1)


def line_graph(n_nodes):
    """Self-only table ``(N, 1)`` with unit weights (reference ``utils.py:282-292``)."""
    table = torch.arange(n_nodes).unsqueeze(1)
    return table, torch.ones_like(table).float()


def get_data_difference(data: torch.Tensor):
    """First difference along time (reference ``utils.py:143-153``)."""
    assert data.ndim == 4, "Data should have 4 dims (B, T, N, C)"
    return data[:, 1:] - data[:, :-1]
