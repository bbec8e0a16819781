"""Graph construction must be bit-exact with the reference (north star): kNN tables, distances,
spatial and temporal weights.  CPU only."""
import contextlib
import io

import pytest
import torch

from _cases import Golden
from _refload import have_reference, load_reference
from mixed_graph_admm_b200 import synth
from mixed_graph_admm_b200 import utils as U

KNN_CASES = ["anchor5", "tiny_f32", "tiny_noexpand", "pems08_f32", "pems04_f32"]


@pytest.mark.parametrize("name", KNN_CASES)
def test_knn_tables_bit_exact_vs_golden(name):
    g = Golden(name)
    k = g.ctor["k"]
    nodes, dists = U.k_nearest_neighbors(g.meta["n_nodes"], g.graph_info["u_edges"], g.graph_info["u_dist"], k)
    assert nodes.dtype == torch.int32 and dists.dtype == torch.float32
    assert torch.equal(nodes.to(torch.int64), g.t("connect_list"))
    assert torch.equal(dists, g.t("dist_list"))


@pytest.mark.parametrize("name", KNN_CASES + ["tiny_physical"])
def test_weights_bit_exact_vs_golden(name):
    g = Golden(name)
    cl, dl = g.t("connect_list"), g.t("dist_list")
    T = g.ctor["T"]
    u = U.undirected_graph_from_distance(cl, dl, u_sigma=g.ctor.get("u_sigma"))
    d = U.directed_graph_from_distance(cl, dl, d_sigma=g.ctor.get("d_sigma"))
    if g.ctor.get("expand_time_dim", True):
        u, d = U.expand_time_dimension(u, T), U.expand_time_dimension(d, T - 1)
    assert torch.equal(u, g.t("u_ew"))
    assert torch.equal(d, g.t("d_ew"))


def test_physical_connect_list_bit_exact():
    g = Golden("tiny_physical")
    cl, dl = U.connect_list(g.meta["n_nodes"], g.graph_info["u_edges"], g.graph_info["u_dist"])
    assert cl.dtype == torch.int64
    assert torch.equal(cl, g.t("connect_list")) and torch.equal(dl, g.t("dist_list"))


def test_python_and_native_knn_agree(libmga):
    from mixed_graph_admm_b200 import _cabi
    for n, ratio, k, iso in [(50, 1.3, 4, True), (400, 1.1, 6, False), (3000, 1.7, 8, True)]:
        gi = synth.road_graph(n, ratio, seed=n, isolate_pair=iso)
        succ = U._adjacency(gi["u_edges"], gi["u_dist"])
        rows = [U._settle_first(succ, s, k + 1) for s in range(n)]
        nn, nd = _cabi.knn_build(n, gi["u_edges"], gi["u_dist"], k)
        for s, row in enumerate(rows):
            m = len(row)
            assert nn[s, :m].tolist() == [a for a, _ in row]
            assert torch.equal(nd[s, :m], torch.tensor([b for _, b in row], dtype=torch.float32))
            assert torch.all(nn[s, m:] == -1) and torch.all(torch.isinf(nd[s, m:]))


def test_knn_duplicate_edges_and_ties(libmga):
    """Repeated (u, v): networkx keeps the first position and the LAST weight; equal distances pop in
    push order.  Both implementations must agree with each other on such input."""
    from mixed_graph_admm_b200 import _cabi
    e = torch.tensor([[0, 1], [0, 2], [0, 1], [1, 0], [2, 0], [1, 3], [3, 1], [2, 3], [3, 2], [3, 4], [4, 3]])
    d = torch.tensor([5., 1., 1., 1., 1., 1., 1., 1., 1., 2., 2.], dtype=torch.float64)
    succ = U._adjacency(e, d)
    assert list(succ[0].items()) == [(1, 1.0), (2, 1.0)]
    nn, nd = _cabi.knn_build(5, e, d, 3)
    for s in range(5):
        row = U._settle_first(succ, s, 4)
        assert nn[s, :len(row)].tolist() == [a for a, _ in row]
    assert nn[0].tolist() == [0, 1, 2, 3]      # tie between 1 and 2 at distance 1: adjacency order


def test_node_without_edges_raises(libmga):
    e = torch.tensor([[0, 1], [1, 0]])
    d = torch.tensor([1., 1.], dtype=torch.float64)
    with pytest.raises(KeyError):
        U.k_nearest_neighbors(3, e, d, 1)


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("n,ratio,k,iso,seed", [(60, 1.3, 4, True, 3), (307, 1.1, 6, False, 11), (500, 1.7, 8, True, 5)])
def test_graph_build_equals_live_reference(n, ratio, k, iso, seed):
    ru, _ = load_reference()
    gi = synth.road_graph(n, ratio, seed=seed, isolate_pair=iso)
    with contextlib.redirect_stdout(io.StringIO()):
        rn, rd = ru.k_nearest_neighbors(n, gi["u_edges"], gi["u_dist"], k)
        cl = rn.to(torch.int64)
        for sig in (None, 50):
            assert torch.equal(ru.undirected_graph_from_distance(cl, rd, u_sigma=sig),
                               U.undirected_graph_from_distance(cl, rd, u_sigma=sig))
            assert torch.equal(ru.directed_graph_from_distance(cl, rd, d_sigma=sig),
                               U.directed_graph_from_distance(cl, rd, d_sigma=sig))
        pa, pb = ru.connect_list(n, gi["u_edges"], gi["u_dist"])
    mn, md = U.k_nearest_neighbors(n, gi["u_edges"], gi["u_dist"], k)
    assert torch.equal(rn, mn) and torch.equal(rd, md)
    qa, qb = U.connect_list(n, gi["u_edges"], gi["u_dist"])
    assert torch.equal(pa, qa) and torch.equal(pb, qb)
