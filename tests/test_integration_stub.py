"""INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.  This test extracts that code
block verbatim, runs it against an object with the reference's attributes (connect_list, u_ew, d_ew, rho, ...) and
checks that it reproduces the drop-in module's result — so the documented binding cannot rot."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "mga_binding.py" in b]
    assert len(stub) == 1, "INTEGRATION.md must hold exactly one mga_binding.py block"
    return stub[0]


def test_stub_mentions_every_argument_of_the_header():
    src = _stub_source()
    header = open(os.path.join(ROOT, "include", "mga.h")).read()
    n_args = len(re.search(r"int mga_admm_solve\((.*?)\);", header, flags=re.S).group(1).split(","))
    argtypes = re.search(r"L\.mga_admm_solve\.argtypes = \[(.*?)\]", src, flags=re.S).group(1)
    assert len([a for a in argtypes.replace("\n", " ").split(",") if a.strip()]) == n_args


@pytest.mark.gpu
def test_documented_stub_reproduces_the_module():
    from mixed_graph_admm_b200 import synth
    from mixed_graph_admm_b200.ADMM import ADMM_algorithm
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)                       # the stub loads the library by its in-tree relative path
    try:
        exec(compile(_stub_source(), "mga_binding.py", "exec"), ns)
        N, k, T, t_in, B = 120, 5, 12, 6, 6
        gi = synth.road_graph(N, 1.4, seed=2)
        blk = ADMM_algorithm(gi, synth.admm_info(N), use_kNN=True, k=k, u_sigma=50, d_sigma=50, t_in=t_in, T=T)
        blk.max_ADMM_iter, blk.max_CG_iter, blk.CG_tol, blk.ADMM_tol = 4, 9, -1.0, -1.0
        y = synth.signals(B, t_in, N, seed=9)
        x_stub = ns["combined_loop_b200"](blk, y)
        x_mod = blk.combined_loop(y, print_info=False)
    finally:
        os.chdir(cwd)
    assert x_stub.shape == x_mod.shape and x_stub.device == y.device
    assert torch.equal(x_stub, x_mod)
