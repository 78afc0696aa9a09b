"""ctypes binding of libdsc_b200.so -- the C-ABI boundary declared in include/dsc_b200.h.

There is deliberately no fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

from .build import LIB_PATH

DTYPE_F16 = 0
DTYPE_BF16 = 1
MAX_KEYS = 80
ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_LAYOUT, ERR_SHAPE = -1, -2, -3, -4

# symbol -> (restype, argtypes); kept in the order of include/dsc_b200.h
PASS_STATS, PASS_FORWARD, PASS_BOTH = 1, 2, 3

SIGNATURES = {
    "dsc_version": (c_int, []),
    "dsc_last_error": (c_char_p, []),
    "dsc_config_set": (c_int, [c_char_p, c_char_p]),
    "dsc_sm_count": (c_int, []),
    "dsc_xattn_workspace_bytes": (c_int, [c_int] * 5 + [POINTER(c_size_t)]),
    "dsc_xattn_stats": (
        c_int,
        [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p]
        + [c_int] * 5 + [c_float, c_int, c_void_p, c_void_p],
    ),
    "dsc_xattn_forward": (
        c_int,
        [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p, c_int,
         c_int, c_void_p, c_float, c_void_p, c_void_p, POINTER(c_int64)]
        + [c_int] * 5 + [c_float, c_int, c_void_p],
    ),
    "dsc_xattn_call_launches": (c_int, [c_int] * 5),
    "dsc_xattn_call_cw": (
        c_int,
        [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p, c_int,
         c_int, c_void_p, c_int, POINTER(ctypes.c_int32), c_void_p, c_float, c_void_p, c_void_p, POINTER(c_int64)]
        + [c_int] * 5 + [c_float, c_int, c_void_p],
    ),
    "dsc_xattn_call_masked": (
        c_int,
        [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p, c_int,
         c_int, c_void_p, POINTER(c_int64), c_void_p, c_float, c_void_p, c_void_p, POINTER(c_int64)]
        + [c_int] * 5 + [c_float, c_int, c_void_p],
    ),
    "dsc_xattn_call": (
        c_int,
        [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p, c_int,
         c_int, c_void_p, c_float, c_void_p, c_void_p, POINTER(c_int64)]
        + [c_int] * 5 + [c_float, c_int, c_void_p],
    ),
    "dsc_xattn_prepared_supported": (c_int, [c_int] * 3),
    "dsc_xattn_kv_image_bytes": (c_int, [c_int] * 4 + [POINTER(c_size_t)]),
    "dsc_xattn_prepare_kv": (
        c_int,
        [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), c_int, POINTER(ctypes.c_int32)]
        + [c_int] * 5 + [c_void_p, c_void_p],
    ),
    "dsc_xattn_call_prepared_launches": (c_int, [c_int] * 5),
    "dsc_xattn_call_prepared": (
        c_int,
        [c_void_p, POINTER(c_int64), c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
         POINTER(c_int64)] + [c_int] * 5 + [c_float, c_int, c_int, c_void_p],
    ),
    "dsc_region_downsample": (c_int, [c_void_p] + [c_int] * 5 + [c_void_p, c_void_p, c_void_p]),
    "dsc_region_accumulate": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
         c_void_p, c_void_p],
    ),
    "dsc_dpmpp2m_step": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double, c_int, c_int,
         c_void_p],
    ),
}


class DscError(RuntimeError):
    """Non-zero return from libdsc_b200 (negative: DSC_ERR_*, positive: cudaError_t)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libdsc_b200 error {code}: {message}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not LIB_PATH.is_file():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no fallback path. "
            "Build it with `python diffusionspatialcontrol_b200/build.py` (needs nvcc; no GPU required)."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def config_set(key: str, value=None) -> None:
    """Kernel-selection override (A/B runs, tests); ``None`` restores the default.  See dsc_config_set."""
    check(lib.dsc_config_set(key.encode(), None if value is None else str(value).encode()))


def check(code: int) -> None:
    if code != 0:
        raise DscError(code, lib.dsc_last_error().decode("utf-8", "replace"))


# ---- optional NVTX ranges around the kernels (K1/K2 attention, K3 region maps, K4 sampler step) --------
NVTX = os.environ.get("DSC_NVTX", "") not in ("", "0")


def nvtx(name: str):
    """Decorator: wraps the call in an NVTX range ``dsc:<name>`` when DSC_NVTX=1 (for ncu --nvtx / nsys); otherwise returns
    the function unchanged (no overhead)."""

    def deco(fn):
        if not NVTX:
            return fn
        import functools

        import torch

        @functools.wraps(fn)
        def wrapped(*a, **kw):
            torch.cuda.nvtx.range_push("dsc:" + name)
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.nvtx.range_pop()

        return wrapped

    return deco
