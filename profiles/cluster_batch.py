"""Cluster mode on BATCHES (float64, fixed counts): one vs two CTAs per SM (MGA_CLUSTER_ONE=1 forces one).
PEMS04-sized graph, T = 12 and T = 24, B = 1024; windows/s through combined_loop on device tensors."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mixed_graph_admm_b200 import synth  # noqa: E402
from mixed_graph_admm_b200.ADMM import ADMM_algorithm  # noqa: E402

N, k, B = 307, 6, 1024
gi = synth.road_graph(N, 1.1, seed=4)
for T in (12, 24):
    y = synth.signals(B, T // 2, N, seed=1, dtype=torch.float64).cuda()
    row = {"T": T, "B": B}
    xs = {}
    for one in ("1", ""):
        if one:
            os.environ["MGA_CLUSTER_ONE"] = one
        else:
            os.environ.pop("MGA_CLUSTER_ONE", None)
        blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=T // 2, T=T)
        blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = 5, 10, -1.0, -1.0
        for _ in range(2):
            x = blk.combined_loop(y, print_info=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            x = blk.combined_loop(y, print_info=False)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        xs[one] = x
        row["one_cta_per_sm" if one else "two_ctas_per_sm"] = round(B / sorted(ts)[2], 1)
    row["bit_identical"] = bool(torch.equal(xs["1"], xs[""]))
    print(json.dumps(row), flush=True)
