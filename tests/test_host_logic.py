"""Host-side logic that needs no GPU: construction, result-list bookkeeping (quirk Q11 types),
sharding + diagnostics reduction over 2 gloo ranks."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _cases import Golden, oracle_from_golden, run_oracle, solver_from_golden
from mixed_graph_admm_b200 import _cabi
from mixed_graph_admm_b200.parallel import reduce_diagnostics, shard_bounds


def _diag_from_trace(tr, B, T, N, with_phi=True, with_zd=True):
    """Partial sums (what the kernels emit) from an oracle trace of B windows."""
    n = tr.outer_iters
    d = np.zeros((n, _cabi.DIAG_COLS))
    dx = np.zeros((n, T, N))
    for i in range(n):
        d[i, _cabi.DIAG_DX2] = tr.x_shift[i] ** 2
        d[i, _cabi.DIAG_X_ZU2] = tr.p_res[i][0] ** 2
        d[i, _cabi.DIAG_DZU2] = tr.d_res[i][0] ** 2
        d[i, _cabi.DIAG_GLR] = tr.glr[i].item() * B
        d[i, _cabi.DIAG_RECOVER2] = tr.recover[i] ** 2
        d[i, _cabi.DIAG_PHI_LDX2] = tr.p_res[i][1] ** 2
        d[i, _cabi.DIAG_DPHI2] = tr.d_res[i][1] ** 2
        d[i, _cabi.DIAG_DGTV] = tr.dgtv[i].item() * B
        d[i, _cabi.DIAG_X_ZD2] = tr.p_res[i][2] ** 2
        d[i, _cabi.DIAG_DZD2] = tr.d_res[i][2] ** 2
        d[i, _cabi.DIAG_DGLR] = tr.dglr[i].item() * B
        dx[i] = tr.dx_mean[i][:, :, 0].numpy() * B
    return d, dx


def test_construction_needs_no_gpu_and_matches_reference_tables():
    g = Golden("tiny_f32")
    blk = solver_from_golden(g)
    assert torch.equal(blk.connect_list, g.t("connect_list"))
    assert torch.equal(blk.u_ew, g.t("u_ew")) and torch.equal(blk.d_ew, g.t("d_ew"))
    assert blk.max_CG_iter == 5 and blk.CG_tol == -1.0
    assert blk.res_name == ['zu', 'phi', 'zd']
    g2 = Golden("tiny_line2")
    b2 = solver_from_golden(g2)
    assert torch.equal(b2.d_ew, g2.t("d_ew")) and torch.equal(b2.time_list, g2.t("time_list"))
    g3 = Golden("tiny_reinit")
    b3 = solver_from_golden(g3)
    assert torch.equal(b3.d_ew, g3.t("d_ew")) and b3.d_ew.dim() == 2        # quirk Q8


def test_compute_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = Golden("tiny_f32")
    blk = solver_from_golden(g)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk.combined_loop(g.y, print_info=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk.apply_op_Lu(g.t("probe_x"))


def test_fill_lists_reproduces_reference_lists_and_types(capsys):
    g = Golden("tiny_f64")
    tr = run_oracle(g)
    B, T, N = g.y.size(0), g.ctor["T"], g.meta["n_nodes"]
    d, dx = _diag_from_trace(tr, B, T, N)
    blk = solver_from_golden(g)
    iters = np.full((tr.outer_iters, 3), -1, dtype=np.int32)
    n_cg = g.limits["max_CG_iter"]
    alpha = torch.rand(tr.outer_iters, 3, n_cg, B, dtype=torch.float64)
    blk._fill_lists(d, dx, iters, tr.outer_iters, alpha, alpha + 1, B, torch.float64, torch.device("cpu"), True)
    np.testing.assert_allclose(blk.x_shift_list, g.z["x_shift"], rtol=1e-12)
    np.testing.assert_allclose(np.array(blk.p_res_list), g.z["p_res"], rtol=1e-12)
    np.testing.assert_allclose(np.array(blk.d_res_list), g.z["d_res"], rtol=1e-12)
    np.testing.assert_allclose(blk.recover_list, g.z["recover"], rtol=1e-12)
    np.testing.assert_allclose([v.item() for v in blk.GLR_list], g.z["glr"], rtol=1e-12)
    np.testing.assert_allclose([v.item() for v in blk.DGTV_list], g.z["dgtv"], rtol=1e-12)
    np.testing.assert_allclose([v.item() for v in blk.DGLR_list], g.z["dglr"], rtol=1e-12)
    np.testing.assert_allclose(torch.stack(blk.delta_x_per_step).numpy(), g.z["delta_x_per_step"], rtol=1e-10)
    # element types the plots rely on (quirk Q11)
    assert isinstance(blk.x_shift_list[0], float) and isinstance(blk.p_res_list[0], list)
    assert isinstance(blk.GLR_list[0], torch.Tensor) and blk.GLR_list[0].dim() == 0
    assert blk.delta_x_per_step[0].shape == (T,)
    assert blk.CG_iter_x == [-1] * tr.outer_iters
    assert isinstance(blk.alpha_x[0], list) and len(blk.alpha_x[0]) == n_cg and blk.alpha_x[0][0].shape == (B,)
    assert torch.equal(blk.beta_zd[1][2], alpha[1, 2, 2] + 1)
    out = capsys.readouterr().out.splitlines()
    assert out[0].startswith("ADMM iters 0: x_CG_iters -1, zu_CG_iters -1, zd_CG_iters -1, pri_err = [")


def test_fill_lists_converged_b1_gives_tensors():
    g = Golden("tiny_tol")
    tr = run_oracle(g)
    T, N = g.ctor["T"], g.meta["n_nodes"]
    d, dx = _diag_from_trace(tr, 1, T, N)
    blk = solver_from_golden(g)
    iters = np.stack([tr.cg_iter_x, tr.cg_iter_zu, tr.cg_iter_zd], 1).astype(np.int32)
    alpha = torch.rand(tr.outer_iters, 3, 100, 1)
    blk._fill_lists(d, dx, iters, tr.outer_iters, alpha, alpha, 1, torch.float32, torch.device("cpu"), False)
    assert blk.CG_iter_x == g.z["cg_iter_x"].tolist()
    assert isinstance(blk.alpha_x[0], torch.Tensor) and blk.alpha_x[0].shape == (tr.cg_iter_x[0],)


def test_strict_quirk_q2_raises_like_the_reference():
    g = Golden("tiny_f32")
    blk = solver_from_golden(g)
    blk.strict_quirks = True
    a = torch.rand(5, 3)
    with pytest.raises(ValueError, match="only one element tensors"):
        blk._coef_lists(a, a, 4, 5, torch.device("cpu"))
    blk.strict_quirks = False
    al, _ = blk._coef_lists(a, a, 4, 5, torch.device("cpu"))
    assert al.shape == (4, 3)


def test_nonfinite_flag_raises_assertion():
    g = Golden("tiny_f32")
    blk = solver_from_golden(g)
    d = np.zeros((1, _cabi.DIAG_COLS))
    d[0, _cabi.DIAG_NONFINITE] = 3
    with pytest.raises(AssertionError, match="NaN"):
        blk._fill_lists(d, np.zeros((1, 6, 24)), np.full((1, 3), -1, np.int32), 1, None, None, 2, torch.float32,
                        torch.device("cpu"), False)


def test_shard_bounds():
    assert shard_bounds(1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for b, w in [(65536, 8), (7, 3), (1, 2)]:
        sb = shard_bounds(b, w)
        assert sb[0][0] == 0 and sb[-1][1] == b and all(a[1] == c[0] for a, c in zip(sb, sb[1:]))


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import admm_oracle as O
        g = Golden("tiny_f64")
        og, prm = oracle_from_golden(g)
        lo, hi = shard_bounds(g.y.size(0), world)[rank]
        T, N = g.ctor["T"], g.meta["n_nodes"]
        tr = O.admm_combined(og, prm, g.y[lo:hi], max_admm_iter=3, max_cg_iter=5, cg_tol=-1.0, admm_tol=-1.0)
        d, dx = _diag_from_trace(tr, hi - lo, T, N)
        gd, gdx, gb = reduce_diagnostics(d, dx, hi - lo)
        blk = solver_from_golden(g)
        blk._fill_lists(gd, gdx, np.full((3, 3), -1, np.int32), 3, None, None, gb, torch.float64,
                        torch.device("cpu"), False)
        xs = [torch.zeros_like(tr.x[:1]).repeat(2, 1, 1, 1) for _ in range(world)]
        pad = torch.zeros_like(xs[0])
        pad[:hi - lo] = tr.x
        dist.all_gather(xs, pad)
        q.put((rank, gb, blk.x_shift_list, blk.p_res_list, [v.item() for v in blk.GLR_list],
               torch.stack(blk.delta_x_per_step).numpy(), torch.cat([xs[0][:2], xs[1][:1]]).numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_reproduces_unsharded_diagnostics():
    """world_size 2 over gloo: each rank solves its slice (oracle stands in for the GPU kernel), the
    partial sums are all-reduced, and both ranks end with the diagnostics of the unsharded batch."""
    g = Golden("tiny_f64")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gb, x_shift, p_res, glr, dxs, x in res:
        assert gb == 3
        np.testing.assert_allclose(x_shift, g.z["x_shift"], rtol=1e-12)
        np.testing.assert_allclose(np.array(p_res), g.z["p_res"], rtol=1e-12)
        np.testing.assert_allclose(glr, g.z["glr"], rtol=1e-12)
        np.testing.assert_allclose(dxs, g.z["delta_x_per_step"], rtol=1e-10)
        np.testing.assert_allclose(x, g.z["x"], rtol=0, atol=1e-13)


def test_solve_sharded_refuses_tolerance_mode(monkeypatch):
    """The reference's stop tests are batch-global (ADMM.py:360, 645): a shard cannot run them on its own windows."""
    from mixed_graph_admm_b200 import parallel
    g = Golden("tiny_f32")
    blk = solver_from_golden(g)
    blk.CG_tol, blk.ADMM_tol = 1e-8, -1.0
    monkeypatch.setattr(parallel.dist, "is_initialized", lambda: True)
    monkeypatch.setattr(parallel.dist, "get_world_size", lambda group=None: 2)
    monkeypatch.setattr(parallel.dist, "get_rank", lambda group=None: 0)
    with pytest.raises(ValueError, match="fixed iteration counts"):
        parallel.solve_sharded(blk, g.y)


def test_plan_cache_pins_the_tensors_behind_its_key():
    """The plan cache is keyed by addresses: the cached entry must keep those tensors alive, or a reassigned d_ew
    could be handed a recycled address and hit a stale plan."""
    import inspect
    from mixed_graph_admm_b200 import ADMM
    src = inspect.getsource(ADMM.ADMM_algorithm._plan)
    assert "key_tensors" in src


def test_knn_pure_python_equals_native(libmga):
    from mixed_graph_admm_b200 import synth, utils
    gi = synth.road_graph(60, 1.3, seed=3, isolate_pair=True)
    a = utils.k_nearest_neighbors(60, gi["u_edges"], gi["u_dist"], 4, native=False)
    b = utils.k_nearest_neighbors(60, gi["u_edges"], gi["u_dist"], 4, native=True)
    assert torch.equal(a[0].to(torch.int64), b[0].to(torch.int64)) and torch.equal(a[1], b[1])


def test_plan_creation_under_address_sanitizer(tmp_path):
    """Every host table mga_plan_create builds (resident schedule, RCM tables, time-tiled tables, row orders) for the
    graphs of the fuzz tests, with mga_plan.cu / mga_schedule.cpp / mga_knn.cpp compiled with AddressSanitizer and the
    CUDA runtime calls answered by host stand-ins (profiles/asan_plan_harness.cpp).  compute-sanitizer is closed on the
    GPU pool; this is the memory check of the host side (it found nothing after the fix of fuzz seed 112, and reports
    that overflow when the fix is taken out)."""
    import glob
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    objs = glob.glob(os.path.join(root, "mixed_graph_admm_b200", "_lib", "mga_resident_ch*_k*.o"))
    if not shutil.which("nvcc") or len(objs) < 12:
        pytest.skip("needs nvcc and the objects of a normal build in _lib/")
    env = dict(os.environ, ASAN_WORK=str(tmp_path))
    r = subprocess.run(["bash", os.path.join(root, "profiles", "asan_plan.sh"), "40", "24"], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "65 plans created and destroyed, 0 refused"      # 40 short (seed 39 among them) + seed 112 + 24 long in r.stdout, r.stdout[-2000:]
    assert "AddressSanitizer" not in r.stderr and "LeakSanitizer" not in r.stderr, r.stderr[-4000:]
