"""Build libexaml_b200.so (CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python -m examl_b200.build [--verbose]

nvcc cross-compiles without a GPU.  The library is git-ignored but travels with the working
tree to the GPU box.  There is no alternative build target: sm_100a only.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libexaml_b200.so")
SOURCES = ["engine.cu", "kernels_dna.cu", "kernels_aa.cu", "kernels_aa_tc.cu", "kernels_sev.cu", "kernels_nr.cu", "kernels_partial.cu"]
HEADERS = ["engine.cuh", "core_tile.cuh", "stage.cuh", "p2p.cuh", "finish.cuh", os.path.join("..", "..", "include", "examl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force=True))
