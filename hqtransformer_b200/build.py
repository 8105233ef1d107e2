"""Builds libhqgraft.so (the sm_100a engine + C ABI) in-tree with nvcc.

    python hqtransformer_b200/build.py [--force] [--verbose]      (run by path: importing the package needs the library)

nvcc cross-compiles for sm_100a without a GPU; the built `.so` is git-ignored but travels to the
GPU box with the repo snapshot.  There is no JIT and no fallback: if the library is missing the
package refuses to work (see _lib.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libhqgraft.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "kernels.cuh", "chain.cuh", "stage1.cuh", "stage1_host.cuh", os.path.join("..", "..", "include", "hqgraft.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, phase_stamps: bool = False) -> str:
    """phase_stamps: the instrumented twin `libhqgraft_phases.so` (-DHQ_PHASE_STAMPS: per-CTA phase stamps of one GEMM launch
    for scripts/gemm_phases.py, loaded with HQ_DEBUG=1 HQGRAFT_LIB=...; never the product library - the stamps cost 6 %)."""
    out = LIB_PATH.replace(".so", "_phases.so") if phase_stamps else LIB_PATH
    if not phase_stamps and not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared"]
    if phase_stamps:
        cmd += ["-DHQ_PHASE_STAMPS"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in SOURCES] + ["-o", out + ".tmp"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libhqgraft.so:\n" + res.stderr[-4000:])
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, phase_stamps="--phase-stamps" in sys.argv))
