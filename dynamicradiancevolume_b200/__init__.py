"""dynamicradiancevolume_b200 — B200-native indirect-lighting path of DynamicRadianceVolume.

``libdrv_gi.so`` (hand-written sm_100a CUDA behind the C-ABI of ``include/drv_gi.h``)
is the product; this package is its Python host mirror of the reference's
``Renderer`` interface for tests and benchmarks.
"""
from . import abi  # noqa: F401
from ._lib import DrvError, load  # noqa: F401
from .renderer import (Camera, Context, IndirectDiffuseMode, Light, Renderer, Scene,  # noqa: F401
                       default_cascade_world_sizes, pack_constant, pack_per_frame, pack_specular, pack_spot_light,
                       pack_volume_info, shard_count, shard_range)

__all__ = ["abi", "load", "DrvError", "Camera", "Light", "Scene", "Renderer", "Context", "IndirectDiffuseMode",
           "pack_constant", "pack_per_frame", "pack_specular", "pack_volume_info", "pack_spot_light", "default_cascade_world_sizes",
           "shard_range", "shard_count"]
