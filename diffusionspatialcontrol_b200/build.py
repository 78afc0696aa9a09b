"""In-tree build of libdsc_b200.so (hand-written sm_100a CUDA behind a C ABI).

``nvcc`` cross-compiles without a GPU; the resulting .so lives next to this file (git-ignored, but it
travels to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# DSC_LIB: alternative library file (debug builds such as -DDSC_TRACE kept next to the product build)
LIB_PATH = Path(os.environ.get("DSC_LIB", str(PKG_DIR / "libdsc_b200.so")))
SOURCES = ["xattn_kernels.cu", "xattn_tc5.cu", "region_kernels.cu", "sampler_kernels.cu", "dsc_capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not LIB_PATH.is_file():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG_DIR.parent / "include" / "dsc_b200.h")
    return any(d.stat().st_mtime > t for d in deps if d.is_file())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libdsc_b200.so; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("DSC_NVCC_EXTRA", "").split()  # e.g. -DDSC_WATCHDOG for a debug build
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
