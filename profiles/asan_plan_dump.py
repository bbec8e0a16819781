"""Dump the graph descriptors the fuzz tests hand to mga_plan_create (both families) for profiles/asan_plan_harness.cpp.
usage: python profiles/asan_plan_dump.py OUT_DIR N_SHORT N_LONG"""
import importlib.util
import os
import struct
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mixed_graph_admm_b200 import ADMM as A, synth  # noqa: E402

spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "tests", "test_gpu_fuzz.py"))
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)
out_dir, n_short, n_long = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.makedirs(out_dir, exist_ok=True)


class _Dumped(Exception):
    pass


_target = [None]


def _fake_plan_init(self, desc, device):
    def arr(v, dt):
        return v.contiguous().to(dt).numpy().tobytes() if isinstance(v, torch.Tensor) else b""
    nu, uw, nd, dw = arr(desc["nbr_u"], torch.int64), arr(desc["u_w"], torch.float32), arr(desc["nbr_d"], torch.int64), arr(desc["d_w"], torch.float32)
    with open(_target[0], "wb") as fh:
        fh.write(struct.pack("9i", desc["n_nodes"], desc["T"], desc["t_in"], desc["ku"], desc["u_w_T"], desc["kd"], desc["d_w_T"],
                             desc["ldrt_mode"], desc["temporal"]))
        fh.write(struct.pack("4q", len(nu) // 8, len(uw) // 4, len(nd) // 8, len(dw) // 4))
        fh.write(nu + uw + nd + dw)
    raise _Dumped()


A._Plan.__init__ = _fake_plan_init
A._device_of = lambda a: torch.device("cpu")
total = 0
for fam, draw, seeds in (("s", fz._draw, list(range(n_short)) + [f for f in fz._FOUND if f >= n_short]),   # (+ the seeds that found bugs)
                         ("l", fz._draw_long, list(range(n_long)))):
    for seed in seeds:
        total += 1
        c = draw(seed)
        N, k = c["N"], min(c["k"], c["N"] - 1)
        gi = synth.road_graph(N, c["ratio"], seed=seed, isolate_pair=N >= 9 and seed % 3 == 0)
        kw = dict(t_in=c["t_in"], T=c["T"], mode=c["mode"])
        if c["variant"] == "knn":
            kw.update(use_kNN=True, k=k, u_sigma=50, d_sigma=50)
        elif c["variant"] == "physical":
            kw.update(use_kNN=False)
        else:
            kw.update(use_kNN=True, k=k, u_sigma=50, use_line_graph=True, skip_connection=c["skip"])
        blk = A.ADMM_algorithm(gi, synth.admm_info(N), **kw)
        if c["varying"]:
            gen = torch.Generator().manual_seed(seed)
            blk.u_ew = blk.u_ew * (0.8 + 0.4 * torch.rand(blk.u_ew.shape, generator=gen))
            blk.d_ew = blk.d_ew * (0.8 + 0.4 * torch.rand(blk.d_ew.shape, generator=gen))
        _target[0] = os.path.join(out_dir, f"{fam}{seed:04d}.bin")
        try:
            blk._plan(c["channels"])
        except _Dumped:
            pass
print("dumped", total, "descriptors to", out_dir)
