"""The oracle against the reference: committed golden vectors, the reference's own known answers,
and (when /root/reference is mounted) the reference run live.  CPU only."""
import numpy as np
import pytest
import torch

from _cases import ALL_CASES, ITERATES, Golden, oracle_from_golden, run_oracle
from _refload import have_reference, run_reference
from oracle import admm_oracle as O


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_reproduces_golden_bit_for_bit(name):
    g = Golden(name)
    tr = run_oracle(g)
    for k in ITERATES:
        if g.has(k):
            assert torch.equal(getattr(tr, k), g.t(k)), f"{name}: {k} differs from the reference"
    assert tr.cg_iter_x == g.z["cg_iter_x"].tolist()
    assert tr.cg_iter_zu == g.z["cg_iter_zu"].tolist()
    if g.has("alpha_zd"):
        assert tr.cg_iter_zd == g.z["cg_iter_zd"].tolist()
    assert np.array_equal(np.array(tr.p_res), g.z["p_res"])
    assert np.array_equal(np.array(tr.d_res), g.z["d_res"])
    assert np.array_equal(np.array(tr.x_shift), g.z["x_shift"])
    assert np.array_equal(np.array(tr.recover), g.z["recover"])
    assert np.array_equal(np.array([v.item() for v in tr.glr]), g.z["glr"])
    assert np.array_equal(np.array([v.item() for v in tr.dgtv]), g.z["dgtv"])
    assert np.array_equal(np.array([v.item() for v in tr.dglr]), g.z["dglr"])
    assert np.array_equal(torch.stack(tr.delta_x_per_step).numpy(), g.z["delta_x_per_step"])


@pytest.mark.parametrize("name", ["anchor5", "tiny_f32", "tiny_f64", "tiny_physical", "tiny_line1", "tiny_line2",
                                  "tiny_noexpand", "pems08_f32"])
def test_oracle_operators_match_golden(name):
    g = Golden(name)
    og, prm = oracle_from_golden(g)
    x, gam = g.t("probe_x"), g.t("probe_gamma")
    assert torch.equal(O.op_lu(og, x), g.t("op_Lu"))
    assert torch.equal(O.op_ldr(og, x), g.t("op_Ldr"))
    assert torch.equal(O.op_ldr_t(og, x), g.t("op_Ldr_T"))
    assert torch.equal(O.op_cldr(og, x), g.t("op_cLdr"))
    assert torch.equal(O.lhs_x(og, prm, x), g.t("op_LHS_x"))
    assert torch.equal(O.lhs_zu(og, prm, x), g.t("op_LHS_zu"))
    assert torch.equal(O.lhs_zd(og, prm, x), g.t("op_LHS_zd"))
    assert torch.equal(O.soft_phi(og, prm, x, gam), g.t("op_phi_direct"))


def test_anchor_matches_survey_hand_values():
    """the 5-node anchor, numbers copied from the survey-time probe of the reference."""
    g = Golden("anchor5")
    assert g.t("connect_list").tolist() == [[0, 1, -1, -1], [1, 0, -1, -1], [2, 3, 4, -1], [3, 2, 4, -1],
                                            [4, 3, 2, -1]]
    og, prm = oracle_from_golden(g)
    x = (torch.arange(15, dtype=torch.float32) / 10).reshape(1, 3, 5, 1)
    np.testing.assert_allclose(O.op_lu(og, x)[0, 0, :, 0], [-.1, .1, -.102613, -.052820, .198092], atol=2e-6)
    ld = O.op_ldr(og, x)[0, :, :, 0]
    assert torch.all(ld[0] == 0)
    np.testing.assert_allclose(ld[1], [.462246, .537754, .463306, .509098, .529673], atol=2e-6)
    lt = O.op_ldr_t(og, x)[0, :, :, 0]
    np.testing.assert_allclose(lt[0], [-.537754, -.462246, -.524351, -.534284, -.441365], atol=2e-6)   # quirk Q1
    np.testing.assert_allclose(lt[2], [1, 1.1, 1.2, 1.3, 1.4], atol=1e-6)
    tr = run_oracle(g)
    np.testing.assert_allclose(tr.x[0, :, :, 0], [[.274671, .327186, .476998, .583516, .575269],
                                                  [.527801, .569191, .731974, .848124, .788992],
                                                  [.752285, .773320, .956733, 1.097053, .942892]], atol=2e-6)


def test_q1_ldr_t_is_transpose_plus_identity_on_t0():
    """apply_op_Ldr_T == L_d^T + diag(1 on the t=0 block)  (quirk Q1, ADMM.py:220-222)."""
    g = Golden("tiny_f64")
    og, _ = oracle_from_golden(g)
    T, N = g.ctor["T"], g.meta["n_nodes"]
    n = T * N
    eye = torch.eye(n, dtype=torch.float64).reshape(n, T, N, 1)
    Ld = O.op_ldr(og, eye).reshape(n, n).T          # columns = images of unit vectors
    LdT = O.op_ldr_t(og, eye).reshape(n, n).T
    diff = LdT - Ld.T
    expect = torch.zeros(n, dtype=torch.float64)
    expect[:N] = 1
    assert torch.allclose(diff, torch.diag(expect), atol=1e-12)
    # and cLdr is exactly L_d^T L_d (symmetric PSD): the Q1 term meets a zero row
    cl = O.op_cldr(og, eye).reshape(n, n).T
    assert torch.allclose(cl, Ld.T @ Ld, atol=1e-12)


def test_band_operator_known_answer_from_notebook():
    """directed_graph.ipynb cells 5-7, 11-12: skip-2 line graph, T=5, L_d [1..5] and L_d^T [1..5]."""
    T, N, skip = 5, 7, 2
    w = torch.ones((N, T, skip))
    w.tril_(diagonal=-1)
    w[:, 0, 0].fill_(1)
    w = w / w.sum(-1, keepdim=True)
    w[:, 0, 0].fill_(0)
    og = O.OracleGraph(nbr=torch.arange(N).unsqueeze(1), u_w=torch.zeros(N, 0), d_w=w.permute(1, 2, 0),
                       line_graph=True, skip=skip,
                       time_list=torch.arange(0, T).unsqueeze(1) - torch.arange(1, skip + 1))
    x = torch.arange(1, 6).float()[None, :, None, None].repeat(1, 1, N, 1)
    assert O.op_ldr(og, x)[0, :, 0, 0].tolist() == [0.0, 1.0, 1.5, 1.5, 1.5]
    assert O.op_ldr_t(og, x)[0, :, 0, 0].tolist() == [-3.5, -1.5, -1.5, 1.5, 5.0]


def test_cg_known_answer_from_cg_script():
    """CG_script.py:49-50: A = [[4,1],[1,3]], b = [1,2] -> x = [1/11, 7/11] in 2 iterations."""
    A = torch.tensor([[4., 1.], [1., 3.]], dtype=torch.float64)
    b = torch.tensor([1., 2.], dtype=torch.float64).reshape(1, 1, 2, 1)
    x, it, al, be = O.cg(lambda v: torch.einsum('ij,btjc->btic', A, v), b, None, max_iter=1000, tol=1e-10)
    assert it == 2
    np.testing.assert_allclose(x.flatten(), [1 / 11, 7 / 11], atol=1e-12)


def test_fixed_iteration_windows_are_independent():
    """In fixed-iteration mode every quantity is per-window: solving windows one at a time gives the
    same x (the property that lets the batch shard over CTAs and GPUs with no collective)."""
    g = Golden("tiny_f64")
    full = run_oracle(g)
    og, prm = oracle_from_golden(g)
    for b in range(g.y.size(0)):
        one = O.admm_combined(og, prm, g.y[b:b + 1], max_admm_iter=3, max_cg_iter=5, cg_tol=-1.0, admm_tol=-1.0)
        assert torch.allclose(one.x, full.x[b:b + 1], rtol=0, atol=1e-13)


@pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("name", ["tiny_f32", "tiny_tol", "tiny_physical", "tiny_line2", "tiny_mask"])
def test_oracle_equals_live_reference(name):
    g = Golden(name)
    ref = run_reference(g.graph_info, g.admm_info, g.y, g.ctor, g.limits, mask=g.mask, init=g.init)
    tr = run_oracle(g)
    for k in ITERATES:
        if k in ref:
            assert torch.equal(getattr(tr, k), ref[k]), k
    assert tr.cg_iter_x == ref["blk"].CG_iter_x


@pytest.mark.parametrize("name", ["two_loops_f32", "two_loops_f64"])
def test_two_loops_oracle_equals_reference_fixture(name):
    """``two_loops`` (ADMM.py:410-508): the oracle restatement is bit-identical to what the live reference left in
    its locals (fixture from tests/golden/make_golden_two_loops.py); the reference appends no residual lists there."""
    from _cases import Golden, oracle_from_golden
    from oracle import admm_oracle as O
    g = Golden(name)
    og, prm = oracle_from_golden(g)
    tr = O.two_loops(og, prm, g.y, max_admm_iter=g.limits["max_ADMM_iter"], max_inner_iter=g.limits["max_inner_iter"],
                     max_cg_iter=g.limits["max_CG_iter"], cg_tol=g.limits["CG_tol"])
    for k in ("x", "zu", "zd", "phi", "gamma", "gamma_u", "gamma_d"):
        assert torch.equal(getattr(tr, k), g.t(k)), k
    assert tr.cg_iter_x == g.z["cg_iter_x"].tolist() and len(tr.cg_iter_zd) == len(tr.cg_iter_x)
    assert g.z["n_lists"].tolist() == [0, 0, 0]
