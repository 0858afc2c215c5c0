"""Build the in-tree native libraries (nvcc cross-compiles sm_100a without a GPU).

  rvtests_b200/librvtests_b200.so   the product: C ABI + CUDA kernels (include/rvtests_b200.h)
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librvtests_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    out = [os.path.join(ROOT, "include", "rvtests_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    if not force and _newer(LIB, srcs):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # box without a toolchain: use the prebuilt library that travelled
        raise RuntimeError("nvcc not found and no prebuilt librvtests_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        os.path.join(CSRC, "rvt_api.cu"), "-o", LIB, "-ldl"]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
