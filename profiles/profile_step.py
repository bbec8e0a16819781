"""Small driver for ncu: a few hot-path steps at PEMS04 shape through the C ABI.

    python profiles/profile_step.py --mode resident --batch 296 --steps 3
    python profiles/profile_step.py --mode streaming --batch 4096 --steps 2
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mixed_graph_admm_b200 import _cabi  # noqa: E402
from mixed_graph_admm_b200.ADMM import _regression_consts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="resident")
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--diag", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
blk, y = bench.build_problem(a.batch, seed=0, device=dev, mode=a.mode)
L = _cabi.lib()
plan, prm = blk._plan(), blk._params()
yd = y.to(dev)
x = torch.empty((a.batch, bench.T_LEN, bench.N_NODES, 1), device=dev)
diag = torch.zeros((bench.N_OUTER, _cabi.DIAG_COLS), dtype=torch.float64, device=dev)
dxs = torch.zeros((bench.N_OUTER, bench.T_LEN, bench.N_NODES), dtype=torch.float64, device=dev)
outs = _cabi.AdmmOutputs()
outs.diag, outs.dx_sum = diag.data_ptr(), dxs.data_ptr()
tm, tv = _regression_consts(bench.T_IN)
st = torch.cuda.current_stream(dev)
for s in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _cabi.check(L.mga_admm_solve(plan.handle, C.byref(prm), _cabi.ptr(yd), bench.T_IN, None, _cabi.ptr(x), a.batch,
                                 0, bench.N_OUTER, bench.N_CG, -1.0, -1.0, tm, tv, a.diag, C.byref(outs),
                                 _cabi.MODE[a.mode], st.cuda_stream))
    e1.record(st)
    torch.cuda.synchronize()
    print(f"step {s}: {e0.elapsed_time(e1):.3f} ms, {a.batch / e0.elapsed_time(e1) * 1e3:.0f} windows/s, "
          f"x checksum {x.double().sum().item():.6f}")
