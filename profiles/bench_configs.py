"""Device-resident timing of the other BASELINE.json configs (not bench lines: context for DESIGN.md).

    python profiles/bench_configs.py [pems08|t288|large20k|pems04_t24 ...] [--mode auto|streaming|resident]
Prints one JSON line per config: windows/s and the algorithmic GB/s of the whole step."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mixed_graph_admm_b200 import _cabi, synth  # noqa: E402
from mixed_graph_admm_b200.ADMM import ADMM_algorithm, _regression_consts  # noqa: E402

CONFIGS = {
    #            N,    k, T,   B,    edge ratio, graph seed
    "pems08": (170, 6, 12, 4096, 1.7, 8),
    "pems04": (307, 6, 12, 8192, 1.1, 4),
    "pems04_t24": (307, 6, 24, 4096, 1.1, 4),
    "t288": (307, 6, 288, 256, 1.1, 4),
    "large20k": (20000, 8, 24, 64, 1.1, 9),
    "pems07_t288": (883, 6, 288, 96, 1.1, 7),      # PEMS07-sized graph, one day at 5 min
    "n600_t96": (600, 6, 96, 512, 1.1, 6),
    "n450_t96": (450, 6, 96, 512, 1.1, 5),
}

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="*", default=["t288", "large20k"])
ap.add_argument("--mode", default="auto")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--diag", type=int, default=1)
ap.add_argument("--batch", type=int, default=0, help="override the config's batch")
a = ap.parse_args()
dev = torch.device("cuda", 0)
L = _cabi.lib()
for name in a.configs:
    N, k, T, B, ratio, gseed = CONFIGS[name]
    B = a.batch or B
    t_in = T // 2
    gi = synth.road_graph(N, ratio, seed=gseed)
    blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T,
                         device=dev, mode=a.mode)
    plan, prm = blk._plan(), blk._params()
    y = synth.signals(B, t_in, N, seed=0).to(dev)
    x = torch.empty((B, T, N, 1), device=dev)
    diag = torch.zeros((bench.N_OUTER, _cabi.DIAG_COLS), dtype=torch.float64, device=dev)
    dxs = torch.zeros((bench.N_OUTER, T, N), dtype=torch.float64, device=dev)
    outs = _cabi.AdmmOutputs()
    outs.diag, outs.dx_sum = diag.data_ptr(), dxs.data_ptr()
    tm, tv = _regression_consts(t_in)
    st = torch.cuda.current_stream(dev)
    l0 = L.mga_launch_count()
    times = []
    for s in range(a.steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        _cabi.check(L.mga_admm_solve(plan.handle, C.byref(prm), _cabi.ptr(y), t_in, None, _cabi.ptr(x), B, 0,
                                     bench.N_OUTER, bench.N_CG, -1.0, -1.0, tm, tv, a.diag, C.byref(outs),
                                     _cabi.MODE[a.mode], st.cuda_stream))
        e1.record(st)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    launches = (L.mga_launch_count() - l0) // (a.steps + 1)
    resident = bool(L.mga_plan_resident_eligible(plan.handle, 0)) and a.mode != "streaming"
    alg = bench.algorithmic_bytes_per_point(t_in=t_in, T=T) * T * N * B
    print(json.dumps({"config": name, "N": N, "k": k, "T": T, "B": B, "mode": "resident" if resident else "streaming",
                      "ms_per_step": ms, "windows_per_s": B / ms * 1e3, "launches_per_step": launches,
                      "algorithmic_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / bench.measured_peaks()[0],
                      "x_checksum": x.double().sum().item()}), flush=True)
    del blk, plan, y, x
    torch.cuda.empty_cache()
