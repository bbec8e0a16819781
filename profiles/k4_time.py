"""Time of one streaming CG solve (10 fixed iterations) at T = 288, N = 307, B = 256 through mga_cg_solve: ms per solve for
the x and z_u systems (layout conversion included).  MGA_LIB / MGA_S4* select variants."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mixed_graph_admm_b200 import _cabi  # noqa: E402

dev = torch.device("cuda", 0)
N, k, T, B = 307, 6, 288, 256
blk, _ = bench.build_problem(1, seed=0, device=dev, mode="streaming", N=N, k=k, T=T, t_in=T // 2)
L = _cabi.lib()
plan, prm = blk._plan(), blk._params()
_cabi.check(L.mga_plan_set_cg_mode(plan.handle, _cabi.MODE["streaming"]))
g = torch.Generator().manual_seed(1)
rhs = torch.rand(B, T, N, 1, generator=g).to(dev)
x = torch.zeros_like(rhs)
st = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for sysname in ("x", "zu"):
    def solve(n_cg):
        _cabi.check(L.mga_cg_solve(plan.handle, _cabi.SYS[sysname], C.byref(prm), _cabi.ptr(rhs), _cabi.ptr(x), None, B, 0,
                                   n_cg, -1.0, None, None, None, st.cuda_stream))
    res = {}
    for n_cg in (10, 30):
        for _ in range(2):
            x.zero_(); solve(n_cg)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            x.zero_(); flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); solve(n_cg); b.record(st); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[n_cg] = best
    per_iter_us = (res[30] - res[10]) / 20 * 1e3
    out.append(f"{sysname}: {per_iter_us:.1f} us/iter (10 its {res[10]:.3f} ms, 30 its {res[30]:.3f} ms)")
print(os.environ.get("MGA_LIB", "default").split("/")[-2] if os.environ.get("MGA_LIB") else "default", " | ".join(out), f"checksum {x.double().sum().item():.4f}")
