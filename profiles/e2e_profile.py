"""Where the end-to-end (host buffers) step spends its time: cProfile over N calls of combined_loop."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
blk, y = bench.build_problem(B, seed=0, device=dev)
yp = y.pin_memory()
for _ in range(3):
    blk.combined_loop(yp, print_info=False)
torch.cuda.synchronize()
n = 20 if B >= 256 else 2
t0 = time.perf_counter()
for _ in range(n):
    blk._reset_lists(all_lists=True)
    blk.combined_loop(yp, print_info=False)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"B={B}: {dt * 1e3:.3f} ms per call, {B / dt:.0f} windows/s")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    blk._reset_lists(all_lists=True)
    blk.combined_loop(yp, print_info=False)
pr.disable()
for knob, val in (("MGA_HOST_PIPE", "0"), ("MGA_HOST_CHUNK", "32"), ("MGA_HOST_CHUNK", "128"), ("MGA_HOST_CHUNK", "256")):
    os.environ[knob] = val
    for _ in range(3):
        blk._reset_lists(all_lists=True)
        blk.combined_loop(yp, print_info=False)
    t0 = time.perf_counter()
    for _ in range(n):
        blk._reset_lists(all_lists=True)
        blk.combined_loop(yp, print_info=False)
    dt = (time.perf_counter() - t0) / n
    print(f"{knob}={val}: {dt * 1e3:.3f} ms per call, {B / dt:.0f} windows/s")
    del os.environ[knob]
blk.keep_cg_coefficients = False
t0 = time.perf_counter()
for _ in range(n):
    blk._reset_lists(all_lists=True)
    blk.combined_loop(yp, print_info=False)
dt = (time.perf_counter() - t0) / n
print(f"without the CG coefficients: {dt * 1e3:.3f} ms per call, {B / dt:.0f} windows/s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
