"""Build libkobato_b200.so (sm_100a only) in-tree with nvcc.

    python kobato-eyes_b200/csrc/build.py [--force] [-v]

Every .cu is compiled to an object under csrc/build/ (in parallel, only when stale) and the objects
are linked into the shared library next to the Python package, so that it travels with the repo
snapshot to the GPU box (git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
OUT = HERE.parent / "kobato_b200" / "libkobato_b200.so"
OBJ = HERE / "build"
SOURCES = ["ke_capi.cu", "ke_multi.cu", "ke_join.cu", "ke_phash.cu", "ke_ssim.cu", "ke_synth.cu", "ke_refine.cu",
           "ke_resize_mma.cu", "ke_scan.cu", "ke_orb.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [HERE / s for s in SOURCES if (HERE / s).exists()]
    headers = list(HERE.glob("*.cuh")) + [ROOT / "include" / "kobato_b200.h"]
    newest_header = max(h.stat().st_mtime for h in headers)
    OBJ.mkdir(exist_ok=True)

    def compile_one(src: Path) -> tuple[Path, bool]:
        obj = OBJ / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, newest_header):
            return obj, False
        cmd = ["nvcc", *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(HERE), "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        print("[build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj, True

    with ThreadPoolExecutor(max_workers=8) as pool:
        done = list(pool.map(compile_one, srcs))
    objs = [o for o, _ in done]
    if force or not OUT.exists() or any(changed for _, changed in done) or \
            any(OUT.stat().st_mtime < o.stat().st_mtime for o in objs):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-o", str(OUT),
               *map(str, objs)]
        print("[build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
