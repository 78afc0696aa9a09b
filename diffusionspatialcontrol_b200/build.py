"""In-tree build of libdsc_b200.so (hand-written sm_100a CUDA behind a C ABI).

``nvcc`` cross-compiles without a GPU; the resulting .so lives next to this file (git-ignored, but it
travels to the GPU box with the repo snapshot).  Every source is compiled to its own object (in parallel, only
when it or a header changed) and the objects are linked into the shared library.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# DSC_LIB: alternative library file (debug builds such as -DDSC_TRACE kept next to the product build)
LIB_PATH = Path(os.environ.get("DSC_LIB", str(PKG_DIR / "libdsc_b200.so")))
OBJ_DIR = PKG_DIR / "build"
SOURCES = ["xattn_kernels.cu", "xattn_tc5.cu", "xattn_x3.cu", "region_kernels.cu", "sampler_kernels.cu",
           "dsc_capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _headers() -> list:
    deps = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG_DIR.parent / "include" / "dsc_b200.h")
    return [d for d in deps if d.is_file()]


def is_stale() -> bool:
    if not LIB_PATH.is_file():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + _headers()
    return any(d.stat().st_mtime > t for d in deps if d.is_file())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libdsc_b200.so; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("DSC_NVCC_EXTRA", "").split()  # e.g. -DDSC_WATCHDOG for a debug build
    nvcc = find_nvcc()
    tag = hashlib.sha1(" ".join(NVCC_FLAGS + extra).encode()).hexdigest()[:8]  # objects of another flag set never mix
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_t = max(h.stat().st_mtime for h in _headers())
    jobs = []
    objs = []
    for s in SOURCES:
        src = CSRC / s
        obj = OBJ_DIR / f"{src.stem}.{tag}.o"
        objs.append(obj)
        if force or not obj.is_file() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_t):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for cmd, proc in ex.map(run, jobs):
            if proc.returncode != 0:
                raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
            if verbose:
                print(" ".join(cmd))
                print(proc.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", str(LIB_PATH), *map(str, objs)]
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
