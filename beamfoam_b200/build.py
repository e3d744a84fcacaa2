"""Builds beamfoam_b200/libbeamgpu.so in-tree with nvcc for sm_100a.

    python -m beamfoam_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "beamgpu.cu")]
HEADERS = [os.path.join(HERE, "csrc", "beam_kernels.cuh"), os.path.join(HERE, "csrc", "beam_long.cuh"),
           os.path.join(HERE, "csrc", "beam_coop.cuh"), os.path.join(HERE, "csrc", "beam_bicgstab.cuh"),
           os.path.join(HERE, "csrc", "beam_contrib.cuh"),
           os.path.join(HERE, "csrc", "beam_math.cuh"),
           os.path.join(ROOT, "include", "beamgpu.h")]
OUTPUT = os.path.join(HERE, "libbeamgpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v",
    # the shared CUDA runtime (libcudart.so.12 is in the loader path of this image, rpath as a second source): the static
    # runtime would embed every runtime entry point's name in the shipped .so, used or not
    "--cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
]


def up_to_date() -> bool:
    if not os.path.exists(OUTPUT):
        return False
    t = os.path.getmtime(OUTPUT)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUTPUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("BEAMGPU_NVCC_EXTRA", "").split()  # e.g. -DBEAM_THREADS=64 -DBEAM_MIN_BLOCKS=5
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", OUTPUT] + SOURCES + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libbeamgpu.so")
    if verbose:
        print(log)
    return OUTPUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(OUTPUT)
