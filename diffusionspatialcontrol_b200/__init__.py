"""B200-native region-masked cross-attention for Stable Diffusion 1.5 (hot path of
duongve13112002/DiffusionSpatialControl), as a drop-in diffusers attention processor backed by
hand-written sm_100a CUDA behind a C ABI (libdsc_b200.so, include/dsc_b200.h)."""
from .attention import compact_region_map, padded_region_map, region_attention, score_stats  # noqa: F401
from .attention_processor import (RegionAttnProcessor, RegionAttnProcessorBaddbmm,  # noqa: F401
                                  RegionIPAdapterAttnProcessor, RegionIPAdapterAttnProcessorBaddbmm,
                                  ip_mask_downsample)
from .region_map import encode_region_map, encode_region_map_sp  # noqa: F401

__all__ = ["RegionAttnProcessor", "RegionAttnProcessorBaddbmm", "region_attention", "score_stats", "encode_region_map", "encode_region_map_sp"]
