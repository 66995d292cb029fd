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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
