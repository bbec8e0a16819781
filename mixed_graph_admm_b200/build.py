"""Build libmga.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libmga.so")
SOURCES = ["mga_plan.cu", "mga_stream.cu", "mga_stream2.cu", "mga_resident.cu", "mga_knn.cpp", "mga_schedule.cpp"]
# the resident kernels are instantiated per (chunks per thread, table slots), one translation unit each
RESIDENT_VARIANTS = [(ch, k) for ch in (1, 2, 3) for k in (4, 6, 8, 10)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmga.so cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False, tag: str = "", defines=()) -> str:
    """Default: _lib/libmga.so.  `tag` + `defines` build an experiment variant (e.g. tag="tab1",
    defines=["MGA_RES_TAB_SMEM=1"]) into _lib/<tag>/libmga.so; select it with MGA_LIB=<path>."""
    out_dir = os.path.join(OUT_DIR, tag) if tag else OUT_DIR
    lib_path = os.path.join(out_dir, "libmga.so")
    os.makedirs(out_dir, exist_ok=True)
    nvcc = _nvcc()
    flags = NVCC_FLAGS + [f"-D{d}" for d in defines]
    headers = [os.path.join(CSRC, "mga_common.cuh"), os.path.join(CSRC, "mga_resident.cuh"), os.path.join(CSRC, "mga_schedule.h"),
               os.path.join(ROOT, "include", "mga.h"), __file__]
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", sp, "-o", obj]
            if src.endswith(".cpp"):
                cmd = [nvcc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
                       "-I", CSRC, "-x", "c++", "-c", sp, "-o", obj]
            jobs.append(cmd)

    inst = os.path.join(CSRC, "mga_resident_inst.cu")
    for tt, k in RESIDENT_VARIANTS:
        obj = os.path.join(out_dir, f"mga_resident_ch{tt}_k{k}.o")
        objs.append(obj)
        if force or _stale(obj, [inst] + headers):
            jobs.append([nvcc] + flags + (["-Xptxas", "-v"] if ptxas_info else []) +
                        [f"-DMGA_CH={tt}", f"-DMGA_K={k}", "-c", inst, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            logs = list(ex.map(run, jobs))
        if ptxas_info:
            for lg in logs:
                sys.stderr.write(lg)
    if jobs or force or _stale(lib_path, objs):
        run([nvcc, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return lib_path


if __name__ == "__main__":
    _tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, ptxas_info="--ptxas" in sys.argv,
                tag=_tag, defines=_defs))
