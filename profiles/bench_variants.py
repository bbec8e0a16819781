"""The variants of the path SURVEY.md §8(f) lists as "next" (line graph / skip connections, mask mode, physical
adjacency, fp64, tolerance mode), timed through the public Python API on device tensors, next to the CPU
oracle (= the reference's torch path, bit-identical) on a bounded sample of the same windows.

    python profiles/bench_variants.py            # one JSON line per variant
    python profiles/bench_variants.py knn_T24_fp64 time_varying_weights     # only these
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mixed_graph_admm_b200 import synth  # noqa: E402
from mixed_graph_admm_b200.ADMM import ADMM_algorithm  # noqa: E402
from oracle import admm_oracle as O  # noqa: E402

N, K, T, T_IN = 307, 6, 24, 12
VARIANTS = [
    # name, ctor kwargs, dtype, batch, cpu sample, fixed (n_outer, n_cg) or None = tolerance mode, mask[, channels]
    ("knn_T24_fp32", dict(use_kNN=True, k=K, u_sigma=50, d_sigma=50), torch.float32, 2048, 16, (5, 10), False),
    ("physical_adjacency", dict(use_kNN=False), torch.float32, 2048, 16, (5, 10), False),
    ("line_graph_skip1", dict(use_kNN=True, k=K, u_sigma=50, use_line_graph=True, skip_connection=1), torch.float32, 2048, 16, (5, 10), False),
    ("line_graph_skip3", dict(use_kNN=True, k=K, u_sigma=50, use_line_graph=True, skip_connection=3), torch.float32, 1024, 16, (5, 10), False),
    ("mask_interpolation", dict(use_kNN=True, k=K, u_sigma=50, d_sigma=50), torch.float32, 1024, 4, (5, 10), True),
    ("two_channels_C2", dict(use_kNN=True, k=K, u_sigma=50, d_sigma=50), torch.float32, 1024, 8, (5, 10), False, 2),
    ("knn_T24_fp64", dict(use_kNN=True, k=K, u_sigma=50, d_sigma=50), torch.float64, 1024, 16, (5, 10), False),
    ("time_varying_weights", dict(use_kNN=True, k=K, u_sigma=50, d_sigma=50), torch.float32, 1024, 8, (5, 10), False, 1, True),
    ("notebook_B1_fp64_tolerance", dict(use_kNN=True, k=4, u_sigma=50, d_sigma=50), torch.float64, 1, 1, None, False),
]

dev = torch.device("cuda", 0)
gi = synth.road_graph(N, 1.1, seed=4)
torch.set_num_threads(os.cpu_count() or 1)
only = set(sys.argv[1:])
for name, kw, dtype, B, b_cpu, fixed, use_mask, *rest in VARIANTS:
    if only and name not in only:
        continue
    Cn = rest[0] if rest else 1
    blk = ADMM_algorithm(gi, synth.admm_info(N), t_in=T_IN, T=T, device=dev, **kw)
    if len(rest) > 1 and rest[1]:       # per-time-step weight tables (T,N,k) / (T-1,N,K): what the unrolling follow-up learns
        gw = torch.Generator().manual_seed(11)
        blk.u_ew = blk.u_ew * (0.8 + 0.4 * torch.rand(blk.u_ew.shape, generator=gw))
        blk.d_ew = blk.d_ew * (0.8 + 0.4 * torch.rand(blk.d_ew.shape, generator=gw))
    if fixed:
        blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = fixed[0], fixed[1], -1.0, -1.0
    else:
        blk.max_ADMM_iter = 30          # bounded: the reference's default 150 takes minutes on the CPU
    if use_mask:
        g = torch.Generator().manual_seed(5)
        full = synth.signals(B, T, N, seed=1, dtype=dtype, smooth=True)
        mask = (torch.rand(B, T, N, 1, generator=g) < 0.6).to(dtype)
        y = full * mask
    else:
        y, mask = synth.signals(B, T_IN, N, seed=1, dtype=dtype, smooth=fixed is None), None
        if Cn > 1:      # C channels: independent signals on the same graph, coupled through the per-window dot products
            y = torch.cat([synth.signals(B, T_IN, N, seed=1 + c, dtype=dtype) for c in range(Cn)], dim=-1).contiguous()
    yd, md = y.to(dev), (mask.to(dev) if mask is not None else None)
    for _ in range(2):
        x = blk.combined_loop(yd, mask=md, print_info=False)
    torch.cuda.synchronize()
    times = []
    for _ in range(15):
        t0 = time.perf_counter()
        x = blk.combined_loop(yd, mask=md, print_info=False)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    gpu_s = sorted(times)[len(times) // 2]          # median: single calls occasionally hit an allocator stall
    # the CPU oracle on the first b_cpu windows
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew, use_knn=bool(kw.get("use_kNN", False)),
                       line_graph=bool(kw.get("use_line_graph", False)), skip=int(kw.get("skip_connection", 1)),
                       time_list=getattr(blk, "time_list", None))
    prm = O.OracleParams(**synth.admm_info(N), t_in=T_IN, T=T)
    ys, ms = y[:b_cpu].contiguous(), (mask[:b_cpu].contiguous() if mask is not None else None)
    t0 = time.perf_counter()
    if use_mask:        # the reference's initial_interpolation only broadcasts for B = 1 (ADMM.py:783-811): one window per call
        xs = []
        for b in range(b_cpu):
            tr = O.admm_combined(og, prm, ys[b:b + 1], mask=ms[b:b + 1], max_admm_iter=int(blk.max_ADMM_iter),
                                 max_cg_iter=int(blk.max_CG_iter), cg_tol=float(blk.CG_tol), admm_tol=float(blk.ADMM_tol))
            xs.append(tr.x)
        x_ref = torch.cat(xs)
    else:
        tr = O.admm_combined(og, prm, ys, mask=ms, max_admm_iter=int(blk.max_ADMM_iter), max_cg_iter=int(blk.max_CG_iter),
                             cg_tol=float(blk.CG_tol), admm_tol=float(blk.ADMM_tol))
        x_ref = tr.x
    cpu_s = time.perf_counter() - t0
    err = ((x[:b_cpu].cpu().double() - x_ref.double()).norm() / x_ref.double().norm()).item()
    print(json.dumps({"variant": name, "dtype": str(dtype).replace("torch.", ""), "batch": B, "path": blk.last_mode,
                      "channels": Cn,
                      "resident": bool(blk._plan(Cn) and __import__("mixed_graph_admm_b200")._cabi.lib().mga_plan_resident_eligible(
                          blk._plan(Cn).handle, 0 if dtype == torch.float32 else 1)) and fixed is not None,
                      "gpu_windows_per_s": B / gpu_s, "gpu_ms": gpu_s * 1e3, "gpu_ms_min": min(times) * 1e3,
                      "cpu_windows_per_s": b_cpu / cpu_s, "cpu_sample": b_cpu, "cpu_threads": torch.get_num_threads(),
                      "speedup": (B / gpu_s) / (b_cpu / cpu_s), "rel_l2_vs_oracle": err,
                      "outer_iters": len(tr.x_shift)}), flush=True)
