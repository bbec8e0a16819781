"""The reference's own call pattern: B = 1, float64, T = 24, t_in = 12, PEMS04-sized graph, tolerance mode (class defaults
CG_tol 1e-8, ADMM_tol 1e-6), 30 outer iterations; solves/s through the public API, cluster kernel vs the general kernels,
with the CPU oracle beside it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mixed_graph_admm_b200 import synth  # noqa: E402
from mixed_graph_admm_b200.ADMM import ADMM_algorithm  # noqa: E402

N, k, T, t_in = 307, 6, 24, 12
gi = synth.road_graph(N, 1.1, seed=4)
out = {}
for dt in (torch.float64, torch.float32):
    y = synth.signals(1, t_in, N, seed=1, smooth=True, dtype=dt)
    for mode in ("auto", "streaming_point"):
        blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T, mode=mode)
        blk.max_ADMM_iter = 30
        yd = y.cuda()
        for _ in range(2):
            blk._reset_lists(all_lists=True)
            x = blk.combined_loop(yd, print_info=False)
        torch.cuda.synchronize()
        n = 10
        t0 = time.perf_counter()
        for _ in range(n):
            blk._reset_lists(all_lists=True)
            x = blk.combined_loop(yd, print_info=False)
        torch.cuda.synchronize()
        dtm = (time.perf_counter() - t0) / n
        out[f"{str(dt)[6:]}_{mode}"] = {"ms_per_solve": 1e3 * dtm, "solves_per_s": 1 / dtm, "cg_iters_x": blk.CG_iter_x[:6],
                                       "outer": len(blk.CG_iter_x), "x_sum": x.double().sum().item()}
print(json.dumps(out))
