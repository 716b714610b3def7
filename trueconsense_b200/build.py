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
    if proc.stderr.strip() and os.environ.get("TC_BUILD_VERBOSE"):
        sys.stderr.write(proc.stderr)


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
    cu = _sources("cuda", (".cu",))
    hdrs = _sources("cuda", (".cuh", ".h")) + [os.path.join(INCLUDE, "trueconsense_b200.h")]
    if force or _newer(CUDA_LIB, cu + hdrs):
        cmd = [nvcc_path(), "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
               "-Xcompiler", "-fvisibility=hidden", "-cudart", "static", "--fmad=false",
               "-I", INCLUDE, "-I", os.path.join(CSRC, "cuda")] + CUDA_ARCH_FLAGS
        if os.environ.get("TC_NVCC_DEFS"):           # e.g. "-DTC_CHUNK_WORDS=6" (tuning experiments)
            cmd += os.environ["TC_NVCC_DEFS"].split()
        if os.environ.get("TC_PTXAS_V"):
            cmd += ["-Xptxas", "-v"]
        cmd += ["-o", CUDA_LIB] + cu + ["-ldl"]
        _run(cmd)
    return CUDA_LIB


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    force = "--force" in sys.argv
    {"host": build_host, "cuda": build_cuda, "all": build_all}[what](force)
    print("built", what)
