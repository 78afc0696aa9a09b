"""The C-ABI library loads on a machine without a GPU, exports every symbol the header declares, and
validates arguments before it touches CUDA (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dsc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from diffusionspatialcontrol_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 9
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/dsc_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_workspace_size():
    from diffusionspatialcontrol_b200 import _lib
    from diffusionspatialcontrol_b200.attention import workspace_bytes

    assert _lib.lib.dsc_version() == 108
    assert workspace_bytes(16, 8, 4096, 40, 77) >= 64 + 16 * 148
    n = ctypes.c_size_t(0)
    assert _lib.lib.dsc_xattn_workspace_bytes(0, 8, 64, 40, 77, ctypes.byref(n)) == _lib.ERR_INVALID_ARGUMENT
    assert _lib.lib.dsc_xattn_workspace_bytes(1, 8, 64, 40, 77, None) == _lib.ERR_INVALID_ARGUMENT


def test_argument_validation_returns_codes_not_crashes():
    from diffusionspatialcontrol_b200 import _lib

    lib = _lib.lib
    I4 = ctypes.c_int64 * 4
    ok_q, ok_k = I4(4096 * 320, 40, 320, 1), I4(77 * 320, 40, 320, 1)
    fake = ctypes.c_void_p(0x1000)  # 16-byte aligned, never dereferenced on these paths
    # unsupported head dim / too many keys / bad dtype / non-positive sizes
    assert lib.dsc_xattn_stats(fake, fake, ok_q, ok_k, None, 2, 8, 4096, 48, 77, 0.1, 0, fake, None) == _lib.ERR_UNSUPPORTED
    assert b"head dim" in lib.dsc_last_error()
    assert lib.dsc_xattn_stats(fake, fake, ok_q, ok_k, None, 2, 8, 4096, 40, 481, 0.1, 0, fake, None) == _lib.ERR_UNSUPPORTED
    assert lib.dsc_xattn_stats(fake, fake, ok_q, ok_k, None, 2, 8, 4096, 40, 77, 0.1, 7, fake, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.dsc_xattn_stats(fake, fake, ok_q, ok_k, None, 2, 8, 0, 40, 77, 0.1, 0, fake, None) == _lib.ERR_INVALID_ARGUMENT
    # additive mask: fp32, 4-byte aligned, non-negative strides; the masked call needs a mask
    assert lib.dsc_xattn_stats(fake, fake, ok_q, ok_k, ctypes.c_void_p(0x1002), 2, 8, 4096, 40, 77, 0.1, 0, fake, None) == _lib.ERR_LAYOUT
    I3m = ctypes.c_int64 * 3
    rc = lib.dsc_xattn_call_masked(fake, fake, fake, ok_q, ok_k, ok_k, fake, 2, 80, fake, I3m(0, -77, 0), None, 1.0, fake, fake,
                                   I3m(4096 * 320, 320, 1), 2, 8, 4096, 40, 77, 0.1, 0, None)
    assert rc == _lib.ERR_LAYOUT and b"mask strides" in lib.dsc_last_error()
    rc = lib.dsc_xattn_call_masked(fake, fake, fake, ok_q, ok_k, ok_k, fake, 2, 80, None, I3m(0, 0, 0), None, 1.0, fake, fake,
                                   I3m(4096 * 320, 320, 1), 2, 8, 4096, 40, 77, 0.1, 0, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT
    # layout contract
    bad = I4(4096 * 320, 4096 * 40, 40, 1)  # contiguous [B,H,L,D]: stride(H) != D
    assert lib.dsc_xattn_stats(fake, fake, bad, ok_k, None, 2, 8, 4096, 40, 77, 0.1, 0, fake, None) == _lib.ERR_LAYOUT
    assert lib.dsc_xattn_stats(ctypes.c_void_p(0x1008), fake, ok_q, ok_k, None, 2, 8, 4096, 40, 77, 0.1, 0, fake, None) == _lib.ERR_LAYOUT
    assert lib.dsc_xattn_stats(None, fake, ok_q, ok_k, None, 2, 8, 4096, 40, 77, 0.1, 0, fake, None) == _lib.ERR_INVALID_ARGUMENT
    # forward: region-map batch must divide the attention batch (the reference raises on the shape mismatch)
    I3 = ctypes.c_int64 * 3
    rc = lib.dsc_xattn_forward(fake, fake, fake, ok_q, ok_k, ok_k, fake, 3, 77, None, 1.0, fake, fake, I3(4096 * 320, 320, 1),
                               2, 8, 4096, 40, 77, 0.1, 0, None)
    assert rc == _lib.ERR_SHAPE and b"Bw=3" in lib.dsc_last_error()
    # sampler step
    assert lib.dsc_dpmpp2m_step(fake, fake, fake, None, 16, 2.0, 0.0, 1.0, 7.5, 1, 0, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.dsc_dpmpp2m_step(fake, fake, fake, None, 16, 1.0, 2.0, 1.0, 7.5, 0, 0, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.dsc_dpmpp2m_step(None, fake, fake, None, 16, 3.0, 2.0, 1.0, 7.5, 0, 0, None) == _lib.ERR_INVALID_ARGUMENT
    # region builder
    assert lib.dsc_region_downsample(None, 2, 512, 512, 64, 64, fake, fake, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.dsc_region_downsample(None, 0, 512, 512, 64, 64, None, None, None) == 0
    assert lib.dsc_region_accumulate(None, None, 0, 0, None, None, None, None, None, 0, 77, fake, None) == _lib.ERR_INVALID_ARGUMENT


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No CPU fallback: importing the binding without the .so raises (checked in a subprocess)."""
    import subprocess
    import sys

    code = (
        "import sys, pathlib; sys.path.insert(0, %r)\n"
        "import importlib.util, types\n"
        "pkg = types.ModuleType('diffusionspatialcontrol_b200'); pkg.__path__ = [%r]\n"
        "sys.modules['diffusionspatialcontrol_b200'] = pkg\n"
        "spec = importlib.util.spec_from_file_location('diffusionspatialcontrol_b200.build', %r)\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "b.LIB_PATH = pathlib.Path(%r) / 'nope.so'\n"
        "sys.modules['diffusionspatialcontrol_b200.build'] = b\n"
        "try:\n"
        "    import diffusionspatialcontrol_b200._lib\n"
        "except ImportError as e:\n"
        "    print('LOUD', e)\n"
    ) % (ROOT, os.path.join(ROOT, "diffusionspatialcontrol_b200"),
         os.path.join(ROOT, "diffusionspatialcontrol_b200", "build.py"), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "LOUD" in out.stdout and "no fallback" in out.stdout, out.stdout + out.stderr
