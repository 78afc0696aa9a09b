"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference hot-path modules from
/root/reference by path, behind a tiny ``diffusers`` stub (diffusers itself is not installed).

Only ``tests/`` and ``scripts/gen_golden.py`` may import this file, and only in the build
container: ``/root/reference`` does not exist on the GPU box, so every caller must check
``reference_available()`` first.  Nothing is copied out of the reference tree; the modules are
executed where they lie.

Reference files loaded:
  source/modules/attention_modify.py            (import list :1-25, processor :414-503)
  source/modules/encode_region_map_function.py  (import list :1-17, builder :21-124)
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DSC_REFERENCE_ROOT", "/root/reference")
_MODULES = os.path.join(REFERENCE_ROOT, "source", "modules")
_cache: dict[str, types.ModuleType] = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_MODULES, "attention_modify.py"))


def _install_diffusers_stub() -> None:
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_dsc_stub", False):
        return  # a real diffusers is importable: use it
    if "diffusers" in sys.modules:
        return

    class _Logger:
        def __getattr__(self, name):
            return lambda *a, **k: None

    class _Logging:
        @staticmethod
        def get_logger(name):
            return _Logger()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._dsc_stub = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    none_names = (
        "_get_model_file delete_adapter_layers is_accelerate_available set_adapter_layers "
        "set_weights_and_activate_adapters scale_lora_layers unscale_lora_layers"
    ).split()
    utils = mod(
        "diffusers.utils",
        USE_PEFT_BACKEND=True,
        logging=_Logging(),
        deprecate=lambda *a, **k: None,
        BaseOutput=object,
        **{n: None for n in none_names},
    )
    emb = mod("diffusers.models.embeddings", ImageProjection=type("ImageProjection", (), {}))
    mu = mod(
        "diffusers.models.modeling_utils",
        _LOW_CPU_MEM_USAGE_DEFAULT=False,
        load_model_dict_into_meta=None,
    )
    models = mod("diffusers.models", embeddings=emb, modeling_utils=mu)
    ip = mod("diffusers.image_processor", IPAdapterMaskProcessor=type("IPAdapterMaskProcessor", (), {}))
    mod("diffusers", utils=utils, models=models, image_processor=ip, DiffusionPipeline=object)


def _load(name: str, filename: str) -> types.ModuleType:
    if name in _cache:
        return _cache[name]
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_diffusers_stub()
    spec = importlib.util.spec_from_file_location(name, os.path.join(_MODULES, filename))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _cache[name] = m
    return m


def attention_modify() -> types.ModuleType:
    return _load("ref_attention_modify", "attention_modify.py")


def encode_region_map_function() -> types.ModuleType:
    return _load("ref_encode_region_map_function", "encode_region_map_function.py")
