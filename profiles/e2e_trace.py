import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
dev = torch.device("cuda", 0)
blk, y = bench.build_problem(1024, seed=0, device=dev)
yp = y.pin_memory()
for _ in range(5):
    blk._reset_lists(all_lists=True); blk.combined_loop(yp, print_info=False)
os.environ["MGA_HOST_TRACE"] = "1"
for _ in range(6):
    blk._reset_lists(all_lists=True)
    t0 = time.perf_counter(); blk.combined_loop(yp, print_info=False); print(f"python call {1e6*(time.perf_counter()-t0):.0f} us", file=sys.stderr)
blk.keep_cg_coefficients = False
for _ in range(3):
    blk._reset_lists(all_lists=True)
    t0 = time.perf_counter(); blk.combined_loop(yp, print_info=False); print(f"python call (no coef) {1e6*(time.perf_counter()-t0):.0f} us", file=sys.stderr)
