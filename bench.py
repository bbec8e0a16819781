#!/usr/bin/env python
"""Benchmark of the Mixed-Graph-ADMM solver hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path (the whole of ``combined_loop``: 5 outer ADMM iterations x
3 CG solves x 10 iterations, diagnostics on) over one batch of synthetic PEMS04-shaped windows
(307 nodes, kNN k=6, T=12, t_in=6, fp32).

  N = 1   BASELINE.json configs[1]: ONE batch of 1024 windows on one B200.
  N > 1   BASELINE.json configs[2]: ONE global batch of 65 536 windows sharded over the N ranks (65 536 / N windows per
          GPU, `parallel.solve_sharded`): no collective on the data path, one NCCL all-reduce of the diagnostics
          partial sums per solve (inside the timed regions).

Prints ONE JSON line (rank 0):
  value      windows/s, whole job, inputs resident in HBM, C-ABI call ``mga_admm_solve`` (+ the diagnostics all-reduce)
  e2e        windows/s through the public API with HOST (pinned) buffers: ``ADMM_algorithm.combined_loop(y_cpu)``
             (N > 1: ``parallel.solve_sharded``) — host->device copy of y, solve, device->host copy of x, the CG
             coefficients and the diagnostics inside the timed region
  roofline   the dominant kernel against the roofline that bounds it (resident kernel: the shared-memory pipe)
  cpu_baseline   the oracle port of the reference's torch-CPU path on the host cores (bounded sample)
  probes     (N = 1) the other BASELINE.json configs on the same clock: T = 288 / B = 256 and 20 000 nodes / k = 8 /
             T = 24 / B = 64 (streaming kernels vs the HBM roofline, each with a 2-window parity sample against the
             oracle), the 65 536-window batch on one GPU, the reference's own notebook call (B = 1, float64, tolerances),
             and the fused CG iteration (``cg_iter.streaming`` at T = 288 where the HBM kernels are the real path,
             ``cg_iter.resident`` / ``cg_iter.pems04_forced_streaming`` at PEMS04 shape)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_NODES, K_NN, T_LEN, T_IN = 307, 6, 12, 6
N_OUTER, N_CG = 5, 10
GLOBAL_BATCH = 65536        # BASELINE.json configs[2]
WORKLOAD = "PEMS04-shaped synthetic: 307 nodes, kNN k=6, T=12, t_in=6, full mixed graph (GLR+DGTV+DGLR), " \
           "5 outer x 3 CG x 10 iters, fp32"
METRIC = "ADMM windows/sec (PEMS04 shape)"
SM_COUNT, SMEM_BYTES_PER_CLK = 148, 128
RESIDENT_COUNTERS = os.path.join("profiles", "r02_resident_counters.json")


def algorithmic_bytes_per_point(n_outer=N_OUTER, n_cg=N_CG, t_in=T_IN, T=T_LEN):
    """Bytes a streaming implementation must move per lattice point for the whole schedule
    (DESIGN.md §5): 48 B per CG iteration for the 2-hop systems, 40 B for z_u."""
    hy = 4.0 * t_in / T
    per_outer = (
        (6 * 4 + hy + 4)                 # RHS_x: gamma, phi, zu, zd, gu, gd, y -> rhs
        + (20 + 12) + 48 * n_cg          # x solve: A x0 (q write+read), r = rhs - A x0; iterations
        + 12 + (8 + 12) + 40 * n_cg      # RHS_zu; z_u solve
        + 12 + (20 + 12) + 48 * n_cg     # RHS_zd; z_d solve
        + 16 + 16 + 16                   # dual ascents, phi prox + gamma
        + (8 * 4 + hy)                   # diagnostics: x, x_old, zu, zu_old, zd, zd_old, phi, phi_old, y
    )
    init = hy + 6 * 4                    # y -> x, zu, zd, phi, (constant duals are not traffic)
    return n_outer * per_outer + init


def cg_iter_bytes_per_point(system="x"):
    return 40.0 if system == "zu" else 48.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        """Median SM clock over the samples taken in [t0, t1] (all samples when the window caught none)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for ts, r in self.rows if len(r) >= 9 and (t0 is None or t0 <= ts <= t1 + 0.06)]
        if not rows:
            rows = [r for _, r in self.rows if len(r) >= 9]
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_problem(batch, seed, device=None, mode="auto", N=N_NODES, k=K_NN, T=T_LEN, t_in=T_IN, ratio=1.1, gseed=4):
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    gi = synth.road_graph(N, ratio, seed=gseed)
    blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T,
                         device=device, mode=mode)
    blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = N_OUTER, N_CG, -1.0, -1.0
    y = synth.signals(batch, t_in, N, seed=seed)
    return blk, y


def oracle_problem(blk, N=N_NODES, T=T_LEN, t_in=T_IN):
    from mixed_graph_admm_b200 import synth
    from oracle import admm_oracle as O
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew)
    prm = O.OracleParams(**synth.admm_info(N), t_in=t_in, T=T)
    return O, og, prm


CPU_BATCH = 32      # windows per call of the CPU path: the batch size at which the reference's torch path is fastest
                    # per window (45 windows/s at B = 32 against 29 at B = 1024 on 8 threads)


def time_oracle(blk, y, budget_s=None):
    """windows/s of the CPU oracle port (all host threads): the windows of `y` in batches of CPU_BATCH, one pass
    (or, with `budget_s`, passes over `y` until that much CPU time is spent).  Returns (windows/s, seconds, windows)."""
    O, og, prm = oracle_problem(blk)
    torch.set_num_threads(os.cpu_count() or 1)
    done, t_all = 0, 0.0
    while True:
        for b0 in range(0, y.size(0), CPU_BATCH):
            ys = y[b0:b0 + CPU_BATCH]
            t0 = time.perf_counter()
            O.admm_combined(og, prm, ys, max_admm_iter=N_OUTER, max_cg_iter=N_CG, cg_tol=-1.0, admm_tol=-1.0)
            t_all += time.perf_counter() - t0
            done += ys.size(0)
            if budget_s is not None and t_all >= budget_s:
                return done / t_all, t_all, done
        if budget_s is None:
            return done / t_all, t_all, done


def loaded_product_libraries():
    """Shared objects of the product mapped into this process (the reference arm must show none)."""
    try:
        with open("/proc/self/maps") as f:
            return sorted({ln.split()[-1] for ln in f if "libmga" in ln})
    except OSError:
        return []


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (its oracle port — the Python reference
    cannot travel to the GPU box) on the host cores; rank 0 only.  Nothing of the product is loaded: the kNN tables
    come from the pure-Python search (`native=False`), the solve is the torch-CPU oracle."""
    if rank != 0:
        return
    sample = 8 * CPU_BATCH          # windows per step (~1.2 s on 16 host threads)
    blk, y = build_problem_cpu(sample)
    for _ in range(min(args.warmup, 1)):
        time_oracle(blk, y[:CPU_BATCH])
    times = []
    for _ in range(args.steps):
        _, dt, _ = time_oracle(blk, y)
        times.append(dt)
    tot = sum(times)
    val = sample * len(times) / tot
    cores = torch.get_num_threads()
    batch = 1024 if world == 1 else GLOBAL_BATCH
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": batch,
                       "sample": f"each step solves {sample} windows of that workload in batches of {CPU_BATCH} (the batch "
                                 "size at which the reference's torch path is fastest per window); windows are "
                                 "independent, so windows/s does not depend on how many of the batch are solved"},
            "same_config": "same workload and metric; the CPU arm solves a bounded sample of the batch per step",
            "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} windows of the workload per step in batches of {CPU_BATCH}, torch-CPU "
                                       f"oracle port (bit-identical to the reference), {cores} threads"},
            "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_libraries_loaded": loaded_product_libraries()}
    print(json.dumps(line), flush=True)


def build_problem_cpu(batch):
    """Graph tables without touching CUDA or libmga.so (for the reference arm)."""
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200 import utils as U

    class Tables:
        pass

    gi = synth.road_graph(N_NODES, 1.1, seed=4)
    t = Tables()
    nodes, dists = U.k_nearest_neighbors(N_NODES, gi["u_edges"], gi["u_dist"], K_NN, native=False)
    t.connect_list = nodes.to(torch.int64)
    t.u_ew = U.expand_time_dimension(U.undirected_graph_from_distance(t.connect_list, dists, u_sigma=50), T_LEN)
    t.d_ew = U.expand_time_dimension(U.directed_graph_from_distance(t.connect_list, dists, d_sigma=50), T_LEN - 1)
    return t, synth.signals(batch, T_IN, N_NODES, seed=0)


class DeviceSolve:
    """``mga_admm_solve`` on device-resident buffers for one problem (the C-ABI call behind ``value`` and the probes)."""

    def __init__(self, blk, y_dev, mode, dev):
        import ctypes as C

        from mixed_graph_admm_b200 import _cabi
        from mixed_graph_admm_b200.ADMM import _regression_consts
        self.C, self.cabi, self.L = C, _cabi, _cabi.lib()
        self.blk, self.y, self.mode, self.dev = blk, y_dev, mode, dev
        self.plan, self.prm = blk._plan(), blk._params()
        B = y_dev.size(0)
        self.B = B
        self.x = torch.empty((B, blk.T, blk.n_nodes, 1), dtype=torch.float32, device=dev)
        self.diag = torch.zeros((N_OUTER * _cabi.DIAG_COLS + N_OUTER * blk.T * blk.n_nodes,), dtype=torch.float64, device=dev)
        self.outs = _cabi.AdmmOutputs()
        self.outs.diag = self.diag.data_ptr()
        self.outs.dx_sum = self.diag.data_ptr() + 8 * N_OUTER * _cabi.DIAG_COLS
        self.tm, self.tv = _regression_consts(blk.t_in)
        self.stream = torch.cuda.current_stream(dev)

    def __call__(self):
        C, c = self.C, self.cabi
        c.check(self.L.mga_admm_solve(self.plan.handle, C.byref(self.prm), c.ptr(self.y), self.blk.t_in, None, c.ptr(self.x),
                                      self.B, c.MGA_F32, N_OUTER, N_CG, -1.0, -1.0, self.tm, self.tv, 1, C.byref(self.outs),
                                      c.MODE[self.mode], self.stream.cuda_stream))


def timed_passes(fn, passes, flush, stream, after=None):
    """Device time of `passes` calls of `fn`, one CUDA-event pair per call on the launching stream, L2 flushed between
    calls (outside the event pairs).  `after` (e.g. the diagnostics all-reduce) runs inside the pair.  Returns ms list."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(passes)]
    for s in range(passes):
        flush.fill_(s & 0xff)
        ev[s][0].record(stream)
        fn()
        if after is not None:
            after()
        ev[s][1].record(stream)
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def probe_config(name, N, k, T, B, ratio, gseed, dev, flush, peak, steps=6):
    """One of the other BASELINE.json configs, device-resident, streaming kernels: windows/s, algorithmic GB/s against the
    HBM roofline, clocks during the run, and parity of the first 2 windows against the CPU oracle."""
    t_in = T // 2
    t0 = time.perf_counter()
    blk, y_host = build_problem(B, seed=0, device=dev, mode="auto", N=N, k=k, T=T, t_in=t_in, ratio=ratio, gseed=gseed)
    y_dev = y_host.to(dev)
    solve = DeviceSolve(blk, y_dev, "auto", dev)
    setup_s = time.perf_counter() - t0
    for _ in range(3):
        solve()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0).start()
    l0 = solve.L.mga_launch_count()
    ta = sampler.mark()
    ms = timed_passes(solve, steps, flush, solve.stream)
    tb = sampler.mark()
    launches = (solve.L.mga_launch_count() - l0) // steps
    clocks = sampler.stop(ta, tb)
    ms_step = sum(ms) / len(ms)
    alg = algorithmic_bytes_per_point(t_in=t_in, T=T) * T * N * B
    gbs = alg / (ms_step * 1e-3) / 1e9
    # parity sample: the first 2 windows against the oracle (the reference's torch-CPU arithmetic)
    from oracle import admm_oracle as O
    _, og, pr = oracle_problem(blk, N=N, T=T, t_in=t_in)
    tr = O.admm_combined(og, pr, y_host[:2].contiguous(), max_admm_iter=N_OUTER, max_cg_iter=N_CG, cg_tol=-1.0, admm_tol=-1.0)
    xs = solve.x[:2].cpu().double()
    err = ((xs - tr.x.double()).norm() / tr.x.double().norm()).item()
    resident = bool(solve.L.mga_plan_resident_eligible(solve.plan.handle, 0))
    out = {"config": name, "N": N, "k": k, "T": T, "t_in": t_in, "batch": B, "kernel_mode": "resident" if resident else "streaming",
           "ms_per_step": ms_step, "windows_per_s": B / (ms_step * 1e-3), "steps": steps, "launches_per_step": int(launches),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "bytes": "algorithmic bytes of the whole step (algorithmic_bytes_per_point) / step time"},
           "parity_vs_oracle_rel_l2": err, "parity_sample": "first 2 windows, x after 5 outer iterations",
           "clocks": clocks, "setup_s": setup_s}
    del solve, blk, y_dev
    torch.cuda.empty_cache()
    return out


def cg_iter_probe(dev, flush, peak, N, k, T, B):
    """The fused CG iteration where the streaming kernels are the real path (windows too long for one CTA): mga_cg_solve
    with 10 and 30 fixed iterations on a batch larger than L2; the difference / 20 is the time of one iteration
    (operator kernel + update kernel), free of the layout conversion and the initial residual."""
    import ctypes as C

    from mixed_graph_admm_b200 import _cabi
    blk, _ = build_problem(1, seed=0, device=dev, mode="streaming", N=N, k=k, T=T, t_in=T // 2)
    L = _cabi.lib()
    plan, prm = blk._plan(), blk._params()
    _cabi.check(L.mga_plan_set_cg_mode(plan.handle, _cabi.MODE["streaming"]))
    g = torch.Generator().manual_seed(1)
    rhs = torch.rand(B, T, N, 1, generator=g).to(dev)
    x = torch.zeros_like(rhs)
    st = torch.cuda.current_stream(dev)
    n = B * T * N
    out = {"N": N, "T": T, "batch": B, "vector_mb": n * 4 / 1e6, "peak": peak, "unit": "GB/s (algorithmic bytes / time)"}
    for sysname in ("x", "zu"):
        res = {}
        for n_cg in (10, 30):
            def solve_cg():
                _cabi.check(L.mga_cg_solve(plan.handle, _cabi.SYS[sysname], C.byref(prm), _cabi.ptr(rhs), _cabi.ptr(x), None, B,
                                           _cabi.MGA_F32, n_cg, -1.0, None, None, None, st.cuda_stream))
            for _ in range(2):
                x.zero_()
                solve_cg()
            torch.cuda.synchronize(dev)
            best = 1e9
            for _ in range(4):
                x.zero_()
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                solve_cg()
                b.record(st)
                torch.cuda.synchronize(dev)
                best = min(best, a.elapsed_time(b))
            res[n_cg] = best
        us = (res[30] - res[10]) / 20 * 1e3
        per_it = cg_iter_bytes_per_point(sysname)
        gbs = per_it * n / (us * 1e-6) / 1e9
        out[sysname] = {"us_per_iter": us, "achieved": gbs, "frac": gbs / peak, "bytes_per_point_per_iter": per_it,
                        "ms_10_iters": res[10], "ms_30_iters": res[30]}
    del rhs, x, blk
    torch.cuda.empty_cache()
    return out


def notebook_probe(dev):
    """The reference's own call pattern (ADMM.py:76-80 defaults): B = 1, float64, T = 24, tolerance mode, 30 outer
    iterations, through the public API with a CPU tensor in and a CPU tensor out; the CPU oracle beside it."""
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    from oracle import admm_oracle as O
    N, k, T, t_in = 307, 6, 24, 12
    gi = synth.road_graph(N, 1.1, seed=4)
    blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, device=dev)
    blk.max_ADMM_iter = 30
    y = synth.signals(1, t_in, N, seed=1, smooth=True, dtype=torch.float64)
    for _ in range(3):
        blk._reset_lists(all_lists=True)
        x = blk.combined_loop(y, print_info=False)
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        blk._reset_lists(all_lists=True)
        x = blk.combined_loop(y, print_info=False)
    dt = (time.perf_counter() - t0) / n
    og = O.OracleGraph(nbr=blk.connect_list, u_w=blk.u_ew, d_w=blk.d_ew)
    prm = O.OracleParams(**synth.admm_info(N), t_in=t_in, T=T)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    tr = O.admm_combined(og, prm, y, max_admm_iter=30)
    t_cpu = time.perf_counter() - t0
    return {"call": "combined_loop(y_cpu) with the class defaults CG_tol 1e-8 / ADMM_tol 1e-6, B = 1, float64, N = 307, T = 24, 30 outer",
            "ms_per_solve": 1e3 * dt, "solves_per_s": 1 / dt, "kernel": "k_admm_cluster (one launch per solve)",
            "cg_iters_x_equal": blk.CG_iter_x == tr.cg_iter_x, "cg_iters_zu_equal": blk.CG_iter_zu == tr.cg_iter_zu,
            "cg_iters_zd_equal": blk.CG_iter_zd == tr.cg_iter_zd, "cg_iters_x": blk.CG_iter_x[:8],
            "rel_l2_vs_oracle": ((x.double() - tr.x.double()).norm() / tr.x.double().norm()).item(),
            "cpu_oracle_ms_per_solve": 1e3 * t_cpu, "cpu_threads": torch.get_num_threads()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="windows per GPU per step (default: 1024 at N=1, 65536/N at N>1)")
    ap.add_argument("--mode", default="auto", choices=["auto", "resident", "streaming"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cg-batch", type=int, default=16384, help="batch of the streaming CG-iteration probe")
    ap.add_argument("--no-cg-probe", action="store_true", help="skip the mga_cg_solve probe (launch lists of the step only)")
    ap.add_argument("--no-probes", action="store_true", help="skip the T=288 / 20k-node / 65536-window probes")
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="inner repeats make the timed region at least this long")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    import ctypes as C

    import torch.distributed as dist
    from mixed_graph_admm_b200 import _cabi, parallel

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(args.warmup, 3)
    # N = 1: configs[1], one batch of 1024.  N > 1: configs[2], ONE batch of 65536 sharded over the ranks.
    if args.batch > 0:
        B, global_batch = args.batch, args.batch * world
    elif world == 1:
        B, global_batch = 1024, 1024
    else:
        lo, hi = parallel.shard_bounds(GLOBAL_BATCH, world)[rank]
        B, global_batch = hi - lo, GLOBAL_BATCH
    blk, y_host = build_problem(B, seed=rank, device=dev, mode=args.mode)     # the global batch = the ranks' shards in rank order
    L = _cabi.lib()
    y_dev = y_host.to(dev)
    y_pin = y_host.pin_memory()
    solve = DeviceSolve(blk, y_dev, args.mode, dev)
    stream = solve.stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def reduce_diag():          # the one collective of the sharded solve: ~74 KB of partial sums (ADMM.py:612-637)
        if world > 1:
            dist.all_reduce(solve.diag, op=dist.ReduceOp.SUM)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value")
    t_probe = []
    for _ in range(warm):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        solve()
        reduce_diag()
        b.record(stream)
        torch.cuda.synchronize(dev)
        t_probe.append(a.elapsed_time(b))
    # inner repeats: every --steps unit is `reps` back-to-back passes so that the timed region is long enough for the
    # clock sampler to see the GPU under this load (the same on every rank: taken from rank 0's estimate)
    est = torch.tensor([min(t_probe) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(est, 0)
    reps = max(1, int(math.ceil(args.min_timed_s / (args.steps * float(est.item())))))
    passes = args.steps * reps
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    l0 = L.mga_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ms = timed_passes(solve, passes, flush, stream, after=reduce_diag)
    barrier()
    t_wall1 = time.perf_counter()
    launches = L.mga_launch_count() - l0
    dev_ms = sum(ms)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    ms_per_step = dev_ms / passes
    value = global_batch * passes / (dev_ms * 1e-3)
    kernel_mode = "resident" if (args.mode != "streaming" and L.mga_plan_resident_eligible(solve.plan.handle, 0)) \
        else "streaming"

    # ---- end to end through the public API with host buffers ("e2e")
    blk.mode = args.mode

    def e2e_call():
        # a caller that solves batch after batch clears the result lists between calls (they grow without bound otherwise,
        # ADMM.py:66-92; here they would also keep every call's pinned alpha / beta blocks alive, and each call would pay
        # for fresh page-locked allocations)
        blk._reset_lists(all_lists=True)
        if world > 1:
            return parallel.solve_sharded(blk, y_pin, y_is_global=False)
        return blk.combined_loop(y_pin, print_info=False)

    t_call = []
    for _ in range(3):      # keeps the previous result alive like the timed loop does (the pinned blocks get cached)
        t1 = time.perf_counter()
        x_host = e2e_call()
        t_call.append(time.perf_counter() - t1)
    est = torch.tensor([min(t_call)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(est, 0)
    e2e_steps = max(args.steps, int(math.ceil(0.6 / float(est.item()))))
    barrier()
    e2e_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        x_host = e2e_call()
        e2e_ms.append(round(1e3 * (time.perf_counter() - t1), 3))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = global_batch * e2e_steps / float(te.item())
    assert x_host.shape == (B, T_LEN, N_NODES, 1) and not x_host.is_cuda and blk.last_mode == "host"
    assert len(blk.alpha_x[-1]) == N_CG and blk.alpha_x[-1][0].shape == (B,)      # the lists of ADMM.py:572-591 are filled
    coef_bytes = 2 * N_OUTER * 3 * N_CG * B * 4

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        npts = T_LEN * N_NODES
        hbm_io = (B * T_IN * N_NODES * 4 + B * npts * 4) / (ms_per_step * 1e-3) / 1e9
        alg_bytes = algorithmic_bytes_per_point() * npts * B
        hbm_equiv = alg_bytes / (ms_per_step * 1e-3) / 1e9
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        cpath = os.path.join(ROOT, RESIDENT_COUNTERS)
        if kernel_mode == "resident" and os.path.exists(cpath):
            # The resident kernel never moves the algorithmic bytes through HBM (y in, x out only): what bounds it is the
            # shared-memory pipe, 128 B per clock and SM.  Wavefronts per window are a property of the kernel and the
            # plan's gather schedule (data-independent); they come from ONE `ncu --set full` capture of this kernel
            # committed under profiles/ (a static file, named here), the rate and the SM clock from this run.
            with open(cpath) as f:
                cnt = json.load(f)
            wf = cnt["smem_wavefronts"] / cnt["batch"]
            achieved = wf * SMEM_BYTES_PER_CLK * (value / world) / 1e9
            smem_peak = SM_COUNT * SMEM_BYTES_PER_CLK * sm_clk * 1e6 / 1e9
            roof = {"bound": "smem", "kernel": "k_admm_resident", "achieved": achieved, "peak": smem_peak, "unit": "GB/s",
                    "frac": achieved / smem_peak,
                    "traffic": (cnt["dram_bytes_read"] + cnt["dram_bytes_write"]) * B / cnt["batch"],
                    "peak_source": f"{SM_COUNT} SMs x {SMEM_BYTES_PER_CLK} B/clk x {sm_clk:.0f} MHz (median SM clock of this run)",
                    "wavefronts_per_window": wf, "wavefronts_source": f"{RESIDENT_COUNTERS} (static: one ncu --set full "
                    f"capture at batch {cnt['batch']}, l1tex__data_pipe_lsu_wavefronts_mem_shared.sum)",
                    "ncu_pipe_pct_of_peak": cnt.get("smem_pipe_pct_of_peak"),
                    "hbm": {"io_GBps": hbm_io, "algorithmic_equivalent_GBps": hbm_equiv, "peak_GBps": peak, "peak_source": peak_src,
                            "note": "real HBM traffic is y in + x out; 'algorithmic_equivalent' = bytes a streaming "
                                    "implementation of the same schedule would move (DESIGN.md §5) / time - context, not a fraction"}}
        else:
            roof = {"bound": "hbm", "kernel": "streaming kernels", "achieved": hbm_equiv, "peak": peak, "unit": "GB/s",
                    "frac": hbm_equiv / peak, "traffic": None, "peak_source": peak_src}
        par = (f"ONE batch of {global_batch} windows sharded over {world} GPUs ({B} per GPU), no data-path collective, one NCCL "
               "all-reduce of the diagnostics per solve") if world > 1 else "one GPU"
        line = {"metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "inner_repeats": reps, "passes_timed": passes,
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": global_batch, "parallelism": par,
                           "kernel_mode": kernel_mode, "diagnostics": "on", "l2": "flushed between timed passes",
                           "baseline_config": "configs[1]" if world == 1 else "configs[2]"},
                "e2e": {"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": int(y_pin.numel() * 4),
                        "d2h_bytes_per_step": int(B * npts * 4 + coef_bytes + N_OUTER * (_cabi.DIAG_COLS + npts) * 8),
                        "steps": e2e_steps, "ms_per_call": e2e_ms[:40], "frac_of_value": e2e_val / value,
                        "api": ("parallel.solve_sharded(blk, y_pinned_cpu_shard) -> " if world > 1 else "") +
                               "ADMM_algorithm.combined_loop(y_pinned_cpu) -> mga_admm_solve_host; x, the CG coefficient arrays "
                               "and the diagnostics sums come back to the host inside the timed call (the python list "
                               "entries over them are created when a list is first read)"},
                "gpu_launches": int(launches), "wall_s_timed_region": t_wall1 - t_wall0, "clocks": clocks, "roofline": roof}

    # ---- the other BASELINE configs on the same clock (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_probes:
        probes = {}
        for name, cfg in (("long_horizon_T288", (307, 6, 288, 256, 1.1, 4)), ("large_graph_20k", (20000, 8, 24, 64, 1.1, 9))):
            try:
                probes[name] = probe_config(name, *cfg, dev=dev, flush=flush, peak=peak)
            except Exception as e:
                probes[name] = {"error": repr(e)[:300]}
        try:        # configs[2]'s batch on ONE GPU: the base of the strong-scaling run
            yb = torch.rand(GLOBAL_BATCH, T_IN, N_NODES, 1, device=dev)
            big = DeviceSolve(blk, yb, args.mode, dev)
            for _ in range(2):
                big()
            torch.cuda.synchronize(dev)
            msb = timed_passes(big, 5, flush, stream)
            probes["global_batch_65536_one_gpu"] = {"batch": GLOBAL_BATCH, "ms_per_step": sum(msb) / len(msb),
                                                    "windows_per_s": GLOBAL_BATCH / (sum(msb) / len(msb) * 1e-3)}
            del big, yb
            torch.cuda.empty_cache()
        except Exception as e:
            probes["global_batch_65536_one_gpu"] = {"error": repr(e)[:300]}
        try:
            probes["notebook_call_b1_f64_tolerance"] = notebook_probe(dev)
        except Exception as e:
            probes["notebook_call_b1_f64_tolerance"] = {"error": repr(e)[:300]}
        line["probes"] = probes

    # ---- the fused CG iteration in streaming mode vs the HBM roofline (vectors larger than L2)
    if rank == 0 and not args.no_cg_probe:
        long_h = None
        try:
            long_h = cg_iter_probe(dev, flush, peak, N=307, k=6, T=288, B=256)
        except Exception as e:
            long_h = {"error": repr(e)[:300]}
        try:
            plan, prm = solve.plan, solve.prm
            Bc = args.cg_batch
            n = Bc * T_LEN * N_NODES
            g = torch.Generator(device="cpu").manual_seed(1)
            rhs = torch.rand(Bc, T_LEN, N_NODES, 1, generator=g).to(dev)
            xw = torch.zeros_like(rhs)
            res = {}
            for impl, cgmode in (("streaming", "streaming"), ("resident", "auto")):
                _cabi.check(L.mga_plan_set_cg_mode(plan.handle, _cabi.MODE[cgmode]))
                res[impl] = {}
                for sysname in ("x", "zu"):
                    def cg():
                        _cabi.check(L.mga_cg_solve(plan.handle, _cabi.SYS[sysname], C.byref(prm), _cabi.ptr(rhs),
                                                   _cabi.ptr(xw), None, Bc, _cabi.MGA_F32, N_CG, -1.0, None, None, None,
                                                   stream.cuda_stream))
                    for _ in range(3):
                        xw.zero_()
                        cg()
                    torch.cuda.synchronize(dev)
                    reps_cg, tot = 5, 0.0
                    for _ in range(reps_cg):
                        xw.zero_()
                        flush.fill_(1)
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(stream)
                        cg()
                        b.record(stream)
                        torch.cuda.synchronize(dev)
                        tot += a.elapsed_time(b)
                    msc = tot / reps_cg
                    per_it = cg_iter_bytes_per_point(sysname)
                    init_b = 32.0 if sysname == "x" else 20.0
                    gbs = (per_it * N_CG + init_b) * n / (msc * 1e-3) / 1e9
                    res[impl][sysname] = {"ms_per_solve": msc, "ms_per_iter": msc / (N_CG + 1), "achieved": gbs,
                                          "frac": gbs / peak, "bytes_per_point_per_iter": per_it}
            _cabi.check(L.mga_plan_set_cg_mode(plan.handle, _cabi.MODE["auto"]))
            # "streaming" = the fused CG iteration where the streaming kernels are the real path (T = 288: windows too long
            # for one CTA; k4_cg + k2_xr, TMA-staged tiles); "resident" = the single-launch solve of windows that fit one CTA
            # (PEMS04 shape); "pems04_forced_streaming" = the PEMS04 shape pushed through the HBM kernels (3-chunk rows)
            line["cg_iter"] = {"streaming": {k2: v for k2, v in long_h.items() if k2 in ("x", "zu", "error")},
                               "streaming_shape": {k2: v for k2, v in long_h.items() if k2 not in ("x", "zu", "error")},
                               "resident": res["resident"], "pems04_forced_streaming": res["streaming"],
                               "batch": Bc, "unit": "GB/s (algorithmic bytes / time)", "peak": peak,
                               "vector_mb": n * 4 / 1e6, "impl": res,
                               "note": "mga_cg_solve, 10 fixed iterations, vectors exceed L2, L2 flushed. 'streaming': fused "
                                       "tile kernels per CG phase, vectors in HBM (the path of windows too large for one "
                                       "CTA), layout conversion included. 'resident': one launch per solve, one window per "
                                       "CTA, rhs/x0 read once - its real HBM traffic is 12 B/pt/solve, so the algorithmic "
                                       "rate exceeds the HBM peak"}
            del rhs, xw
        except Exception as e:  # the headline line must still be printed
            line["cg_iter"] = {"error": str(e)[:200]}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only; bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        time_oracle(blk, y_host[:CPU_BATCH])                       # warm-up pass
        val, dt, n_done = time_oracle(blk, y_host, budget_s=12.0)
        cores = torch.get_num_threads()
        # and parity of the timed kernel on a sample of those windows
        from oracle import admm_oracle as O
        _, og, pr = oracle_problem(blk)
        ys = y_host[:CPU_BATCH].contiguous()
        tr = O.admm_combined(og, pr, ys, max_admm_iter=N_OUTER, max_cg_iter=N_CG, cg_tol=-1.0, admm_tol=-1.0)
        err = ((x_host[:CPU_BATCH].double() - tr.x.double()).norm() / tr.x.double().norm()).item()
        line["cpu_baseline"] = {"value": val, "unit": "windows/s", "cores": cores, "kind": "port",
                                "sample": f"{n_done} windows of the step's batch in batches of {CPU_BATCH} ({dt:.1f} s of CPU "
                                          f"time), torch-CPU oracle port (bit-identical to the reference), {cores} threads"}
        line["parity_vs_oracle_rel_l2"] = err
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
