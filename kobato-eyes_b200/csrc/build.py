"""Build libkobato_b200.so (sm_100a only) in-tree with nvcc.

    python kobato-eyes_b200/csrc/build.py [--force]

The shared library lands next to the Python package so that it travels with the repo snapshot
to the GPU box (git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
OUT = HERE.parent / "kobato_b200" / "libkobato_b200.so"
SOURCES = ["ke_capi.cu", "ke_join.cu", "ke_phash.cu", "ke_ssim.cu", "ke_synth.cu", "ke_refine.cu", "ke_resize_mma.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
    "--shared", "-cudart", "static",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [HERE / s for s in SOURCES if (HERE / s).exists()]
    deps = srcs + list(HERE.glob("*.cuh")) + [ROOT / "include" / "kobato_b200.h"]
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(HERE), "-o", str(OUT), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
