"""Full-size parity of the BASELINE.json configs, the host-buffer entry point, the fused elementwise exports and
the NCCL-sharded solve (round-2 items of VERDICT.md).  All through the C ABI, all against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from _cases import Golden, max_rel, oracle_from_golden, rel_err, solver_from_golden

pytestmark = pytest.mark.gpu


def _problem(N, k, T, B, ratio, gseed, yseed, n_outer=5, n_cg=10, mode="auto"):
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    from oracle import admm_oracle as O
    t_in = T // 2
    gi = synth.road_graph(N, ratio, seed=gseed)
    blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, mode=mode)
    blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = n_outer, n_cg, -1.0, -1.0
    y = synth.signals(B, t_in, N, seed=yseed)
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew)
    prm = O.OracleParams(**synth.admm_info(N), t_in=t_in, T=T)
    return blk, y, og, prm


def _check_against_oracle(blk, y, og, prm, n_outer, n_cg):
    from oracle import admm_oracle as O
    blk.keep_iterates = True
    x = blk.combined_loop(y.cuda(), print_info=False).cpu()
    tr = O.admm_combined(og, prm, y, max_admm_iter=n_outer, max_cg_iter=n_cg, cg_tol=-1.0, admm_tol=-1.0)
    assert rel_err(x, tr.x) <= 1e-5 and max_rel(x, tr.x) <= 2e-5, rel_err(x, tr.x)
    its = {k: v.cpu() for k, v in blk.last_iterates.items()}
    assert rel_err(its["zu"], tr.zu) <= 1e-5 and rel_err(its["zd"], tr.zd) <= 1e-5
    ldx = O.op_ldr(og, tr.x).double().norm().item()
    assert (its["phi"].double() - tr.phi.double()).norm().item() <= 1e-5 * max(tr.phi.double().norm().item(), ldx)
    for name, ref in (("gamma", tr.gamma), ("gamma_u", tr.gamma_u), ("gamma_d", tr.gamma_d)):
        assert rel_err(its[name], ref) <= 1e-4
    # Residual norms over ~1e6 lattice points: the reference's own fp32 `norm()` is up to 1.9e-4 away from its fp64 run
    # here (20 000 nodes, primal[phi] of the first iteration: 31.8521 in fp32, 31.8581 in fp64; the kernels, which add
    # the per-CTA partial sums in double, give 31.8581).  So the lists are held to the fp64 run of the same arithmetic.
    tr64 = O.admm_combined(og, prm, y.double(), max_admm_iter=n_outer, max_cg_iter=n_cg, cg_tol=-1.0, admm_tol=-1.0)
    np.testing.assert_allclose(blk.x_shift_list, tr64.x_shift, rtol=5e-5)
    np.testing.assert_allclose(np.array(blk.p_res_list), np.array(tr64.p_res), rtol=5e-5, atol=1e-7)
    np.testing.assert_allclose(np.array(blk.d_res_list), np.array(tr64.d_res), rtol=5e-5, atol=1e-7)
    nz = (tr.phi != 0).float().mean().item()
    assert 0.02 < nz < 0.98, "the DGTV prox must be exercised (SURVEY.md §8d)"
    return x


def test_long_horizon_t288_full_schedule():
    """BASELINE.json configs[3]: 307 nodes, T = 288, the full 5 outer x 10 CG schedule, iterates and residual lists."""
    blk, y, og, prm = _problem(307, 6, 288, 2, 1.1, 4, 1)
    _check_against_oracle(blk, y, og, prm, 5, 10)
    assert blk.last_mode == "device"


def test_large_graph_20k_nodes_full_schedule():
    """BASELINE.json configs[4] at its full width: 20 000 nodes, kNN k = 8, T = 24, 5 outer x 10 CG, B = 2."""
    from mixed_graph_admm_b200 import _cabi
    blk, y, og, prm = _problem(20000, 8, 24, 2, 1.1, 9, 2)
    assert _cabi.lib().mga_plan_resident_eligible(blk._plan().handle, 0) == 0
    _check_against_oracle(blk, y, og, prm, 5, 10)


def test_full_batch_t288_windows_are_independent():
    """configs[3] at B = 256: every window of the full batch equals the same window solved in a batch of 2
    (size-independent property; the B = 2 solve is checked against the oracle above)."""
    blk, y, _, _ = _problem(307, 6, 288, 256, 1.1, 4, 1)
    yd = y.cuda()
    full = blk.combined_loop(yd, print_info=False)
    for lo in (0, 131, 254):
        part = blk.combined_loop(yd[lo:lo + 2].contiguous(), print_info=False)
        assert rel_err(part, full[lo:lo + 2]) <= 1e-6       # per-window dots: double atomics, free order


# ---- host-buffer entry point (ADVICE: scratch overlap of concurrent chunk launches; CG coefficients on the CPU path)
@pytest.mark.parametrize("B", [1001, 7, 300])
@pytest.mark.parametrize("pipe", ["1", "0"])
def test_host_entry_equals_device_entry_odd_batches(B, pipe, monkeypatch):
    """combined_loop(y_cpu) == combined_loop(y_cuda) bit for bit at batch sizes whose chunks are unequal (1001 ->
    251+251+251+248 in the chunked mode, where the last launch's parking scratch used to overlap its predecessor's)
    and smaller than one chunk; MGA_HOST_PIPE=1: one persistent launch fed chunk by chunk, 0: chunked launches."""
    from test_gpu_parity import _pems04
    monkeypatch.setenv("MGA_HOST_PIPE", pipe)
    blk, y = _pems04(B, seed=3)
    blk.mode = "resident"
    xd = blk.combined_loop(y.cuda(), print_info=False).cpu()
    dev_lists = (list(blk.x_shift_list), [v.item() for v in blk.GLR_list])
    a_dev = torch.stack(list(blk.alpha_zd[-1]))
    for pinned in (False, True):
        xh = blk.combined_loop(y.pin_memory() if pinned else y, print_info=False)
        assert blk.last_mode == "host" and not xh.is_cuda
        assert torch.equal(xh, xd), (B, pipe, pinned, rel_err(xh, xd))
        np.testing.assert_allclose(blk.x_shift_list[-5:], dev_lists[0], rtol=1e-6)
        np.testing.assert_allclose([v.item() for v in blk.GLR_list][-5:], dev_lists[1], rtol=1e-6)
        # the CG coefficient lists of ADMM.py:572-591: n_cg tensors of shape (B,) per solve, equal to the device path's
        assert len(blk.alpha_x) % 5 == 0 and isinstance(blk.alpha_x[-1], list) and len(blk.alpha_x[-1]) == 10
        assert blk.alpha_x[-1][0].shape == (B,) and not blk.alpha_x[-1][0].is_cuda
        assert torch.equal(torch.stack(list(blk.alpha_zd[-1])), a_dev.cpu())
        blk.init_iterations('None')
        blk.d_ew = Golden("pems04_f32").t("d_ew")          # init_iterations resets d_ew (quirk Q8); restore


def test_host_entry_streaming_mode_returns_coefficients():
    """The chunked host entry of the streaming kernels (windows too long for one CTA) also fills alpha / beta."""
    blk, y, og, prm = _problem(64, 4, 40, 9, 1.3, 3, 5, n_outer=2, n_cg=6)
    xd = blk.combined_loop(y.cuda(), print_info=False).cpu()
    a_dev = torch.stack(list(blk.alpha_x[0])).cpu()
    xh = blk.combined_loop(y, print_info=False)
    assert blk.last_mode == "host"
    assert rel_err(xh, xd) <= 1e-6
    a_host = torch.stack(list(blk.alpha_x[2]))
    assert a_host.shape == (6, 9)
    np.testing.assert_allclose(a_host.numpy(), a_dev.numpy(), rtol=1e-5)


# ---- the three fused elementwise exports, called directly (they were exported but never exercised)
def _dev(t):
    return t.cuda().contiguous()


@pytest.mark.parametrize("name", ["tiny_f32", "tiny_f64", "pems08_f32", "tiny_line2", "tiny_physical"])
def test_rhs_x_dual_ascent_prox_exports_against_oracle(name):
    from mixed_graph_admm_b200 import _cabi
    from oracle import admm_oracle as O
    g = Golden(name)
    blk = solver_from_golden(g)
    og, prm = oracle_from_golden(g)
    L = _cabi.lib()
    plan, p = blk._plan(), blk._params()
    dt = g.dtype
    tol = 2e-6 if dt == torch.float32 else 1e-13
    B, T, N, t_in = 3, g.ctor["T"], g.meta["n_nodes"], g.ctor["t_in"]
    gen = torch.Generator().manual_seed(17)
    r = lambda *s: torch.randn(*s, generator=gen, dtype=dt)    # noqa: E731
    gam, phi, zu, zd, gu, gd, x = (r(B, T, N, 1) for _ in range(7))
    y = torch.rand(B, t_in, N, 1, generator=gen, dtype=dt)
    st = torch.cuda.current_stream().cuda_stream
    # RHS_x (ADMM.py:552-559)
    out = torch.empty(B, T, N, 1, dtype=dt, device="cuda")
    d = [_dev(t) for t in (gam, phi, zu, zd, gu, gd, y)]
    _cabi.check(L.mga_rhs_x(plan.handle, C.byref(p), *[_cabi.ptr(t) for t in d], t_in, _cabi.ptr(out), B, _cabi.dtype_id(dt), st))
    hty = torch.cat([y, torch.zeros(B, T - t_in, N, 1, dtype=dt)], 1)
    ref = O.op_ldr_t(og, gam + prm.rho * phi) / 2 + (prm.rho_u * zu + prm.rho_d * zd) / 2 - (gu + gd) / 2 + hty
    assert rel_err(out.cpu(), ref) <= tol, (name, rel_err(out.cpu(), ref))
    # dual ascent (ADMM.py:595-597)
    gz, xd, zud = _dev(gu), _dev(x), _dev(zu)          # (named: a temporary would be freed, and its memory reused, before the launch)
    _cabi.check(L.mga_dual_ascent(plan.handle, float(prm.rho_u), _cabi.ptr(xd), _cabi.ptr(zud), _cabi.ptr(gz), B,
                                  _cabi.dtype_id(dt), st))
    assert rel_err(gz.cpu(), gu + prm.rho_u * (x - zu)) <= tol
    # phi prox + gamma ascent (ADMM.py:401-408, 603-605)
    gm, ph = _dev(gam), torch.empty(B, T, N, 1, dtype=dt, device="cuda")
    _cabi.check(L.mga_prox_phi_dual(plan.handle, C.byref(p), _cabi.ptr(xd), _cabi.ptr(gm), _cabi.ptr(ph), B,
                                    _cabi.dtype_id(dt), st))
    ldx = O.op_ldr(og, x)
    ph_ref = O.soft_phi(og, prm, x, gam)
    assert rel_err(ph.cpu(), ph_ref) <= 10 * tol
    assert rel_err(gm.cpu(), gam + prm.rho * (ph_ref - ldx)) <= 10 * tol


# ---- BASELINE.json configs[2]: ONE batch sharded over NCCL ranks == the unsharded solve
def _nccl_rank(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from mixed_graph_admm_b200 import parallel
    from test_gpu_parity import _pems04
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        blk, y = _pems04(301, seed=5)
        blk.device = f"cuda:{rank}"
        res = {}
        for where in ("cpu", "cuda"):
            yy = y if where == "cpu" else y.cuda()
            x = parallel.solve_sharded(blk, yy, gather=True)
            res[where] = (x.cpu(), list(blk.x_shift_list[-5:]), [list(v) for v in blk.p_res_list[-5:]],
                          [v.item() for v in blk.GLR_list[-5:]], torch.stack([v.cpu() for v in blk.delta_x_per_step[-5:]]),
                          blk.alpha_x[-1][0].shape[0])
        blk.CG_tol = 1e-8
        try:
            parallel.solve_sharded(blk, y)
            rejected = False
        except ValueError:
            rejected = True
        q.put((rank, res, rejected))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_nccl_two_rank_sharded_solve_equals_unsharded():
    import torch.multiprocessing as mp
    from test_gpu_parity import _pems04
    blk, y = _pems04(301, seed=5)
    x_ref = blk.combined_loop(y.cuda(), print_info=False).cpu()
    ref = (list(blk.x_shift_list), [list(v) for v in blk.p_res_list], [v.item() for v in blk.GLR_list],
           torch.stack([v.cpu() for v in blk.delta_x_per_step]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res, rejected in out:
        assert rejected, "tolerance mode must be refused: its stop tests are batch-global (ADMM.py:360, 645)"
        for where, (x, x_shift, p_res, glr, dxs, nloc) in res.items():
            # shards of 151 + 150 windows, all-gathered (the shorter shard is padded for the collective, the pad dropped)
            assert x.shape == x_ref.shape and torch.equal(x, x_ref), (rank, where)
            np.testing.assert_allclose(x_shift, ref[0], rtol=1e-6)
            np.testing.assert_allclose(np.array(p_res), np.array(ref[1]), rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(glr, ref[2], rtol=1e-6)
            np.testing.assert_allclose(dxs.numpy(), ref[3].numpy(), rtol=1e-5, atol=1e-9)
            assert nloc in (150, 151)


# ---- cluster mode: one thread-block cluster per window, the signal's own precision, stop tests on the device
@pytest.mark.parametrize("name", ["tiny_f64", "tiny_c2_f64", "tiny_mask_f64", "tiny_f32", "tiny_prox", "tiny_physical", "tiny_line1", "pems08_f32"])
def test_cluster_mode_fixed_iterations_match_reference(name):
    """Fixed iteration counts through the cluster kernel (float64 takes it by default; float32 goldens are sent there by
    asking for tolerance -1 with a batch the kernel also serves): every iterate and list against the golden."""
    from mixed_graph_admm_b200 import _cabi
    from test_gpu_parity import DUAL_TOL, TOL, _check_iterates, _check_lists
    g = Golden(name)
    blk = solver_from_golden(g)
    blk.keep_iterates = True
    L = _cabi.lib()
    y = g.y.cuda()
    if g.dtype == torch.float64:
        l0 = L.mga_launch_count()
        x = blk.combined_loop(y, mask=None if g.mask is None else g.mask.cuda(), print_info=False).cpu()
        assert L.mga_launch_count() - l0 == 1, "float64 with fixed counts is one cluster launch"
    else:
        # float32 + fixed counts belongs to the resident kernel; reach the cluster kernel through the C ABI
        import ctypes as C
        from mixed_graph_admm_b200.ADMM import _regression_consts
        plan, prm = blk._plan(g.y.size(-1)), blk._params()
        B, T, N = g.y.size(0), g.ctor["T"], g.meta["n_nodes"] * g.y.size(-1)
        outs = _cabi.AdmmOutputs()
        its = {k: torch.zeros(B, T, N, 1, device="cuda") for k in ("zu", "zd", "phi", "gamma", "gamma_u", "gamma_d")}
        for k, v in its.items():
            setattr(outs, k, v.data_ptr())
        xd = torch.empty(B, T, N, 1, device="cuda")
        tm, tv = _regression_consts(g.ctor["t_in"])
        rc = L.mga_cluster_solve(plan.handle, C.byref(prm), _cabi.ptr(y.contiguous()), _cabi.ptr(xd), B, 0, g.limits["max_ADMM_iter"],
                                 g.limits["max_CG_iter"], -1.0, -1.0, tm, tv, 0, C.byref(outs), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc)
        torch.cuda.synchronize()
        x = xd.reshape(g.t("x").shape).cpu()
        blk.last_iterates = {k: v.reshape(g.t("x").shape) for k, v in its.items()}
    blk.last_iterates = {k: v.cpu() for k, v in blk.last_iterates.items()}
    _check_iterates(blk, x, g, TOL[g.dtype])
    if g.dtype == torch.float64:
        _check_lists(blk, g, 1e-10)


def test_cluster_mode_notebook_call_is_one_launch():
    """The reference's own call pattern (B = 1, float64, T = 24, tolerances): one launch, CG counts of the golden."""
    from mixed_graph_admm_b200 import _cabi
    g = Golden("pems04_t24_tol_f64")
    blk = solver_from_golden(g)
    l0 = _cabi.lib().mga_launch_count()
    x = blk.combined_loop(g.y, print_info=False)
    assert _cabi.lib().mga_launch_count() - l0 == 1
    assert blk.CG_iter_x == g.z["cg_iter_x"].tolist() and blk.CG_iter_zu == g.z["cg_iter_zu"].tolist()
    assert blk.CG_iter_zd == g.z["cg_iter_zd"].tolist()
    assert rel_err(x, g.t("x")) <= 1e-11


@pytest.mark.parametrize("dtype,varying,T", [(torch.float64, False, 12), (torch.float64, False, 24), (torch.float32, True, 24),
                                              (torch.float64, True, 12)])
def test_cluster_mode_batches_two_ctas_per_sm(monkeypatch, dtype, varying, T):
    """A batch that oversubscribes the GPU with clusters: the two-CTAs-per-SM instantiation with compact tables (16-bit
    indices, in-list weights through slot indices; per-step weight slices when the tables vary in time) against the
    one-per-SM one (MGA_CLUSTER_ONE=1; same arithmetic up to fma contraction) and against the oracle on the first windows."""
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    from oracle import admm_oracle as O
    N, k, B = 170, 6, 48
    gi = synth.road_graph(N, 1.7, seed=8)
    y = synth.signals(B, T // 2, N, seed=3, dtype=dtype)
    gen = torch.Generator().manual_seed(21)
    xs = {}
    for one in ("1", None):
        if one:
            monkeypatch.setenv("MGA_CLUSTER_ONE", one)
        else:
            monkeypatch.delenv("MGA_CLUSTER_ONE", raising=False)
        blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=T // 2, T=T)
        if varying:
            if one:
                fu = 0.8 + 0.4 * torch.rand(blk.u_ew.shape, generator=gen)
                fd = 0.8 + 0.4 * torch.rand(blk.d_ew.shape, generator=gen)
            blk.u_ew, blk.d_ew = blk.u_ew * fu, blk.d_ew * fd
        blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = 4, 8, -1.0, -1.0
        l0 = _launches()
        xs[one] = blk.combined_loop(y.cuda(), print_info=False).cpu()
        assert blk.last_mode == "device" and _launches() - l0 == 1           # one cluster launch for the whole batch
    f64 = dtype == torch.float64
    assert rel_err(xs[None], xs["1"]) <= (1e-14 if f64 else 2e-6)
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew)
    prm = O.OracleParams(**synth.admm_info(N), t_in=T // 2, T=T)
    tr = O.admm_combined(og, prm, y[:3], max_admm_iter=4, max_cg_iter=8, cg_tol=-1.0, admm_tol=-1.0)
    assert rel_err(xs[None][:3], tr.x) <= (1e-12 if f64 else 1e-5)


def _launches():
    from mixed_graph_admm_b200 import _cabi
    return _cabi.lib().mga_launch_count()


def test_cluster_mode_time_varying_weights_tolerance_call():
    """Per-time-step weight tables (T,N,k) / (T-1,N,K) with the notebooks' call pattern (B = 1, float64,
    tolerances): one cluster launch, the oracle's CG counts and iterates."""
    from mixed_graph_admm_b200 import _cabi
    from oracle import admm_oracle as O
    g = Golden("pems04_t24_tol_f64")
    blk = solver_from_golden(g)
    gen = torch.Generator().manual_seed(5)
    blk.u_ew = blk.u_ew * (0.7 + 0.6 * torch.rand(blk.u_ew.shape, generator=gen, dtype=blk.u_ew.dtype))
    blk.d_ew = blk.d_ew * (0.7 + 0.6 * torch.rand(blk.d_ew.shape, generator=gen, dtype=blk.d_ew.dtype))
    assert blk.u_ew.dim() == 3 and blk.d_ew.dim() == 3
    l0 = _cabi.lib().mga_launch_count()
    x = blk.combined_loop(g.y, print_info=False)
    assert _cabi.lib().mga_launch_count() - l0 == 1
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew)
    prm = O.OracleParams(**g.admm_info, t_in=g.ctor["t_in"], T=g.ctor["T"])
    tr = O.admm_combined(og, prm, g.y, max_admm_iter=blk.max_ADMM_iter, max_cg_iter=blk.max_CG_iter, cg_tol=blk.CG_tol,
                         admm_tol=blk.ADMM_tol)
    assert blk.CG_tol > 0 and blk.ADMM_tol > 0
    assert blk.CG_iter_x == list(tr.cg_iter_x) and blk.CG_iter_zu == list(tr.cg_iter_zu) and blk.CG_iter_zd == list(tr.cg_iter_zd)
    assert rel_err(x, tr.x) <= 1e-11


@pytest.mark.parametrize("name", ["two_loops_f32", "two_loops_f64"])
def test_two_loops_matches_reference(name):
    """``two_loops`` (ADMM.py:410-508): returns nothing like the reference, appends only the CG lists, and
    leaves in ``last_iterates`` what the reference's locals hold at the end (fixture captured from the live reference)."""
    from test_gpu_parity import DUAL_TOL, TOL
    g = Golden(name)
    blk = solver_from_golden(g)
    assert blk.two_loops(g.y) is None
    n_solves = g.limits["max_ADMM_iter"] * g.limits["max_inner_iter"]
    assert blk.CG_iter_x == g.z["cg_iter_x"].tolist() and len(blk.CG_iter_zu) == n_solves and len(blk.CG_iter_zd) == n_solves
    assert len(blk.alpha_x) == n_solves and isinstance(blk.alpha_x[0], list) and len(blk.alpha_x[0]) == g.limits["max_CG_iter"]
    assert blk.p_res_list == [] and blk.x_shift_list == [] and blk.GLR_list == []       # no residual lists in two_loops
    its = blk.last_iterates
    for k in ("x", "zu", "zd"):
        assert not its[k].is_cuda and rel_err(its[k], g.t(k)) <= TOL[g.dtype], (k, rel_err(its[k], g.t(k)))
    for k in ("gamma", "gamma_u", "gamma_d"):
        assert rel_err(its[k], g.t(k)) <= DUAL_TOL[g.dtype], (k, rel_err(its[k], g.t(k)))
    og, _ = oracle_from_golden(g)
    from oracle import admm_oracle as O
    ldx = O.op_ldr(og, g.t("x")).double().norm().item()
    assert (its["phi"].double() - g.t("phi").double()).norm().item() <= TOL[g.dtype] * max(g.t("phi").double().norm().item(), ldx)
    a0 = torch.stack(list(blk.alpha_x[0])).cpu()[:2].double().numpy()
    np.testing.assert_allclose(a0, g.z["alpha_x"][0, :2], rtol=1e-4 if g.dtype == torch.float32 else 1e-9)


def test_host_entry_survives_blocking_launches():
    """The pipelined host launch waits for its uploads on the device, so every upload must be queued before the launch call:
    with CUDA_LAUNCH_BLOCKING=1 (or under a profiler that serialises launches) the launch call only returns when the
    kernel is over.  Run the host entry that way in a fresh process; it must finish and agree with the device entry."""
    import subprocess
    code = (
        "import sys, time, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_gpu_parity import _pems04\n"
        "blk, y = _pems04(600, seed=9)\n"
        "t0 = time.time(); xh = blk.combined_loop(y.pin_memory(), print_info=False); dt = time.time() - t0\n"
        "xd = blk.combined_loop(y.cuda(), print_info=False).cpu()\n"
        "assert blk.last_mode == 'device' and torch.equal(xh, xd), 'host entry differs'\n"
        "assert dt < 3.0, dt\n"
        "print('ok', dt)\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
