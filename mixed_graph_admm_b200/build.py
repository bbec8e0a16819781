"""Build libmga.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libmga.so")
SOURCES = ["mga_plan.cu", "mga_stream.cu", "mga_stream2.cu", "mga_resident.cu", "mga_cluster.cu", "mga_knn.cpp", "mga_schedule.cpp"]
# the resident kernels are instantiated per (chunks per thread, table slots), one translation unit each
RESIDENT_VARIANTS = [(ch, k) for ch in (1, 2, 3) for k in (4, 6, 8, 10)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmga.so cannot be built")
    return exe


def _digest(paths, extra=()):
    """Content hash of the files a target is built from (+ the flags): staleness does not depend on mtimes, which a
    snapshot copy to the GPU box need not preserve, and the objects need not travel with the library."""
    h = hashlib.sha1()
    for x in extra:
        h.update(str(x).encode())
        h.update(b"\0")
    for f in paths:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


# -lineinfo on every translation unit (ncu's source page); the embedded PTX text it brings is 2/3 of an object, so the
# instantiations for windows of T <= 8 (CH = 1, 2: test-sized problems nobody profiles) get it only with MGA_LINEINFO_ALL=1
def _lineinfo(ch):
    return ch >= 3 or os.environ.get("MGA_LINEINFO_ALL") == "1"


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False, tag: str = "", defines=()) -> str:
    """Default: _lib/libmga.so.  `tag` + `defines` build an experiment variant (e.g. tag="tab1",
    defines=["MGA_RES_TAB_SMEM=1"]) into _lib/<tag>/libmga.so; select it with MGA_LIB=<path>.
    _lib/build_manifest.json records the content hash every object and the library were built from (what decides
    staleness) and _lib/build_log.json what the last call compiled."""
    out_dir = os.path.join(OUT_DIR, tag) if tag else OUT_DIR
    lib_path = os.path.join(out_dir, "libmga.so")
    os.makedirs(out_dir, exist_ok=True)
    flags = NVCC_FLAGS + [f"-D{d}" for d in defines]
    headers = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".h"))] + \
              [os.path.join(ROOT, "include", "mga.h")]
    man_path = os.path.join(out_dir, "build_manifest.json")
    try:
        with open(man_path) as fh:
            manifest = json.load(fh)
    except Exception:
        manifest = {}
    units = []      # (object, source, extra flags, lineinfo)
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if os.path.exists(sp):
            units.append((os.path.join(out_dir, os.path.splitext(src)[0] + ".o"), sp, [], True))
    inst = os.path.join(CSRC, "mga_resident_inst.cu")
    for tt, k in RESIDENT_VARIANTS:
        units.append((os.path.join(out_dir, f"mga_resident_ch{tt}_k{k}.o"), inst, [f"-DMGA_CH={tt}", f"-DMGA_K={k}"], _lineinfo(tt)))
    want = {os.path.basename(o): _digest([sp] + headers, flags + extra + [li]) for o, sp, extra, li in units}
    lib_key = _digest([], sorted(want.items()))
    log = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "library": lib_path, "compiled": [], "linked": False}
    if not force and os.path.exists(lib_path) and manifest.get("libmga.so") == lib_key:
        log["note"] = "library is up to date with the sources (content hash): nothing compiled"
        with open(os.path.join(out_dir, "build_log.json"), "w") as fh:
            json.dump(log, fh, indent=1)
        return lib_path
    nvcc = _nvcc()
    jobs = []
    for o, sp, extra, li in units:
        name = os.path.basename(o)
        if force or not os.path.exists(o) or manifest.get(name) != want[name]:
            fl = [f for f in flags if li or f != "-lineinfo"]
            cmd = [nvcc] + fl + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", sp, "-o", o]
            if sp.endswith(".cpp"):
                cmd = [nvcc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
                       "-I", CSRC, "-x", "c++", "-c", sp, "-o", o]
            jobs.append((name, cmd))

    def run(job):
        name, cmd = job
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
        for name, _ in jobs:
            manifest[name] = want[name]
            log["compiled"].append(name)
    run(("libmga.so", [nvcc, "-shared", "-o", lib_path] + [o for o, _, _, _ in units] + ["-gencode", "arch=compute_100a,code=sm_100a"]))
    manifest["libmga.so"] = lib_key
    log["linked"] = True
    with open(man_path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    with open(os.path.join(out_dir, "build_log.json"), "w") as fh:
        json.dump(log, fh, indent=1)
    return lib_path


if __name__ == "__main__":
    _tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, ptxas_info="--ptxas" in sys.argv,
                tag=_tag, defines=_defs))
