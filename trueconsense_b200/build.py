"""In-tree builds of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

* ``libtchost.so``  — host C: BGZF/BAM io + synthetic reads (gcc, zlib, OpenMP)
* ``libtcb200.so``  — the CUDA hot path behind the C-ABI of include/trueconsense_b200.h
                      (nvcc, sm_100a only)
``python -m trueconsense_b200.build [host|cuda|all]``  (the oracle under oracle/ builds itself:
``oracle.pileup.build()``; nothing here touches it)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "trueconsense_b200")
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")

HOST_LIB = os.path.join(PKG, "libtchost.so")
CUDA_LIB = os.path.join(PKG, "libtcb200.so")

CUDA_ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"build failed: {cmd[0]} exited {proc.returncode}")
    if (proc.stderr.strip() or proc.stdout.strip()) and (os.environ.get("TC_BUILD_VERBOSE") or os.environ.get("TC_PTXAS_V")):
        sys.stderr.write(proc.stdout + proc.stderr)


def _sources(subdir: str, exts: tuple[str, ...]) -> list[str]:
    d = os.path.join(CSRC, subdir)
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_host(force: bool = False) -> str:
    srcs = _sources("host", (".c",))
    deps = srcs + [os.path.join(INCLUDE, "tc_host.h")]
    if force or _newer(HOST_LIB, deps):
        _run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wno-unused-result",
              "-I", INCLUDE, "-o", HOST_LIB] + srcs + ["-lz", "-lm"])
    return HOST_LIB


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False) -> str:
    """Every .cu is compiled to its own object (in parallel, only when it or a header changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor

    cu = _sources("cuda", (".cu",))
    hdrs = _sources("cuda", (".cuh", ".h")) + [os.path.join(INCLUDE, "trueconsense_b200.h")]
    objdir = os.path.join(ROOT, "build", "cuda")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
             "--fmad=false", "-I", INCLUDE, "-I", os.path.join(CSRC, "cuda")] + CUDA_ARCH_FLAGS
    if os.environ.get("TC_NVCC_DEFS"):           # e.g. "-DTC_FLAT_CW=6" (tuning experiments)
        flags += os.environ["TC_NVCC_DEFS"].split()
        force = True
    if os.environ.get("TC_PTXAS_V"):
        flags += ["-Xptxas", "-v"]
        force = True
    objs, jobs = [], []
    for src in cu:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([nvcc_path()] + flags + ["-c", "-o", obj, src])
    stale = [o for o in os.listdir(objdir) if o.endswith(".o") and os.path.join(objdir, o) not in objs]
    for o in stale:
        os.remove(os.path.join(objdir, o))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(_run, jobs))
    if jobs or stale or not os.path.exists(CUDA_LIB) or _newer(CUDA_LIB, objs):
        _run([nvcc_path(), "-shared", "-cudart", "static"] + CUDA_ARCH_FLAGS + ["-o", CUDA_LIB] + objs + ["-ldl"])
    return CUDA_LIB


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    force = "--force" in sys.argv
    {"host": build_host, "cuda": build_cuda, "all": build_all}[what](force)
    print("built", what)
