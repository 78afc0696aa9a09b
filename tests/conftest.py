import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The C-ABI library must exist before the package can be imported (there is no fallback path).
    spec = importlib.util.spec_from_file_location("_dsc_build", os.path.join(ROOT, "diffusionspatialcontrol_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except RuntimeError as e:  # no nvcc on this machine: use the prebuilt .so if it is there
        if not mod.LIB_PATH.is_file():
            raise pytest.UsageError(f"cannot build libdsc_b200.so: {e}")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def dsc_config():
    """Kernel-selection overrides through the C ABI (dsc_config_set), restored to the defaults afterwards.  The library
    reads the DSC_* environment variables only once, at first use, so tests switch families through this call."""
    from diffusionspatialcontrol_b200 import _lib

    touched = []

    def set_(key, value):
        touched.append(key)
        _lib.config_set(key, value)

    yield set_
    for key in touched:
        _lib.config_set(key, None)
