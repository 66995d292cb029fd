"""Build libian_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python neural-photo-editor_b200/build.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["ian_api.cu", "tapgemm_simt.cu", "tapgemm_tc.cu", "decout_tc.cu", "conv1_tc.cu", "edge_kernels.cu"]
HDR = ["tapgemm.h", "edge.h", "tc_ptx.cuh", "../../include/ian_b200.h"]
LIB = os.path.join(HERE, "libian_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, "csrc", f) for f in SRC + HDR] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(HERE, "csrc", f) for f in SRC]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


C_HOST_SRC = os.path.join(HERE, "..", "examples", "c_host", "ian_cli.c")
C_HOST_BIN = os.path.join(HERE, "..", "examples", "c_host", "ian_cli")


def build_c_host(force: bool = False) -> str:
    """the plain-C host program over the C-ABI (examples/c_host): strict C99 against include/ian_b200.h, linked to
    the in-tree library with an $ORIGIN-relative rpath so it runs from the snapshot on the GPU box."""
    lib = build()
    deps = [C_HOST_SRC, os.path.join(HERE, "..", "include", "ian_b200.h"), lib]
    if not force and os.path.exists(C_HOST_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(C_HOST_BIN) for d in deps):
        return C_HOST_BIN
    cmd = [os.environ.get("CC", "gcc"), "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic",
           "-I", os.path.join(HERE, "..", "include"), C_HOST_SRC, "-o", C_HOST_BIN,
           "-L", HERE, "-lian_b200", "-Wl,-rpath,$ORIGIN/../../" + os.path.basename(HERE)]
    subprocess.run(cmd, check=True)
    return C_HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_c_host(force="--force" in sys.argv))
