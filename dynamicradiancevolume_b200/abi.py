"""ctypes images of the POD structs declared in ``include/drv_gi.h``.

Byte-exact std140 / std430 layouts of the reference's uniform blocks and
buffers (``shader/globalubos.glsl``, ``shader/lightcache.glsl``); the sizes are
asserted at import time so a drift between this file and the header fails
loudly.
"""
import ctypes as C

import numpy as np

DRV_MAX_CASCADES = 4
DRV_MAX_LIGHTS = 16
DRV_IPC_HANDLE_BYTES = 64

DRV_OK = 0
DRV_ERR_INVALID = -1
DRV_ERR_CUDA = -2
DRV_ERR_CAPACITY = -3
DRV_ERR_NOT_BOUND = -4
DRV_ERR_NO_DEVICE = -5
DRV_ERR_PEER = -6

DRV_VOXELIZE_CLEAR = 1
DRV_VOXELIZE_FINISH = 2
DRV_HDR_RGBA16F_ADD = 0
DRV_HDR_RGBA32F_WRITE = 1
DRV_HDR_RGBA16F_WRITE = 2
DRV_FRAME_PREPARE_RSM = 1
DRV_FRAME_GRAPH = 2
DRV_FRAME_APPLY_OWN_ROWS = 4
DRV_FRAME_GATHER_IMAGE = 8
DRV_FRAME_VOXELIZE = 16

STAGE_NAMES = ["VoxelizeScene", "VoxelBlendMipMap", "AllocateCaches", "LightCaches", "ApplyCaches",
               "PrepareRSM", "GatherKernel", "ConeKernel"]

f32 = C.c_float
i32 = C.c_int32
u32 = C.c_uint32


class Constant(C.Structure):
    """``Constant`` block, globalubos.glsl:2-30 (80 bytes)."""
    _fields_ = [
        ("ShCosLobeFactor0", f32), ("ShCosLobeFactor1", f32), ("ShCosLobeFactor2n2_p1_n1", f32),
        ("ShCosLobeFactor20", f32), ("ShCosLobeFactor2p2", f32),
        ("ShEvaFactor0", f32), ("ShEvaFactor1", f32), ("ShEvaFactor2n2_p1_n1", f32),
        ("ShEvaFactor20", f32), ("ShEvaFactor2p2", f32),
        ("BackbufferResolution", i32 * 2),
        ("VoxelResolution", i32), ("AddressVolumeResolution", i32), ("NumAddressVolumeCascades", i32),
        ("MaxNumLightCaches", u32),
        ("SpecularEnvmapTotalSize", i32), ("SpecularEnvmapPerCacheSize_Texel", i32),
        ("SpecularEnvmapPerCacheSize_Texcoord", f32), ("SpecularEnvmapNumCachesPerDimension", i32),
    ]


class PerFrame(C.Structure):
    """``PerFrame`` block, globalubos.glsl:33-43 (288 bytes)."""
    _fields_ = [
        ("Projection", f32 * 16), ("ViewProjection", f32 * 16), ("InverseView", f32 * 16),
        ("InverseViewProjection", f32 * 16),
        ("CameraPosition", f32 * 3), ("_pad0", f32),
        ("CameraDirection", f32 * 3), ("PassedTime", f32),
    ]


class CAVCascade(C.Structure):
    """``CAVCascade``, globalubos.glsl:48-62 (64 bytes)."""
    _fields_ = [
        ("Min", f32 * 3), ("WorldVoxelSize", f32), ("Max", f32 * 3), ("_padding0", f32),
        ("DecisionMin", f32 * 3), ("_padding1", f32), ("DecisionMax", f32 * 3), ("_padding2", f32),
    ]


class VolumeInfo(C.Structure):
    """``VolumeInfo`` block, globalubos.glsl:65-79 (288 bytes)."""
    _fields_ = [
        ("VolumeWorldMin", f32 * 3), ("VoxelSizeInWorld", f32),
        ("VolumeWorldMax", f32 * 3), ("CAVTransitionZoneSize", f32),
        ("AddressVolumeCascades", CAVCascade * DRV_MAX_CASCADES),
    ]


class SpotLight(C.Structure):
    """``SpotLight`` block, globalubos.glsl:88-112 (224 bytes)."""
    _fields_ = [
        ("LightIntensity", f32 * 3), ("ShadowNormalOffset", f32),
        ("ShadowBias", f32), ("_pad0", f32 * 3),
        ("LightPosition", f32 * 3), ("_pad1", f32),
        ("LightDirection", f32 * 3), ("LightCosHalfAngle", f32),
        ("LightViewProjection", f32 * 16), ("InverseLightViewProjection", f32 * 16),
        ("RSMRenderResolution", i32), ("RSMReadResolution", i32), ("ValAreaFactor", f32),
        ("IndirectShadowComputationLod", f32), ("IndirectShadowComputationBlockSize", f32),
        ("IndirectShadowComputationSampleInterval", i32),
        ("IndirectShadowComputationSuperValWidth", f32), ("IndirectShadowSamplingOffset", f32),
    ]


class CacheCounter(C.Structure):
    """``LightCacheCounter``, lightcache.glsl:83-90 (16 bytes)."""
    _fields_ = [("NumCacheLightingThreadGroupsX", u32), ("NumCacheLightingThreadGroupsY", u32),
                ("NumCacheLightingThreadGroupsZ", u32), ("TotalLightCacheCount", i32)]


class Config(C.Structure):
    """``drv_config`` (drv_gi.h)."""
    _fields_ = [
        ("max_cache_count", u32), ("cav_cascades", u32), ("cav_resolution", u32),
        ("voxel_resolution", u32), ("sh_order", u32), ("indirect_shadow", u32),
        ("cascade_transitions", u32), ("backbuffer_width", u32), ("backbuffer_height", u32),
        ("max_lights", u32), ("max_rsm_resolution", u32), ("device", i32),
        ("stream", C.c_void_p), ("gather_variant", u32), ("indirect_specular", u32),
        ("specular_per_cache_size", u32), ("specular_fill_holes_level", u32),
    ]


class Buffers(C.Structure):
    """``drv_buffers`` (drv_gi.h)."""
    _fields_ = [
        ("entries", C.c_void_p), ("entry_stride", u32), ("max_cache_count", u32),
        ("counter", C.c_void_p), ("cav_atlas", C.c_void_p),
        ("cav_width", u32), ("cav_height", u32), ("cav_depth", u32),
        ("voxel_chain", C.c_void_p), ("voxel_target", C.c_void_p),
        ("voxel_resolution", u32), ("voxel_levels", u32), ("voxel_chain_bytes", C.c_uint64),
        ("vpls", C.c_void_p * DRV_MAX_LIGHTS), ("shadow_blocks", C.c_void_p * DRV_MAX_LIGHTS),
        ("rsm_flux_mips", C.c_void_p * DRV_MAX_LIGHTS), ("rsm_normal_mips", C.c_void_p * DRV_MAX_LIGHTS),
        ("rsm_depth_mips", C.c_void_p * DRV_MAX_LIGHTS),
        ("hdr16", C.c_void_p),
        ("rsm_flux0", C.c_void_p * DRV_MAX_LIGHTS), ("rsm_normal0", C.c_void_p * DRV_MAX_LIGHTS),
        ("rsm_depth0", C.c_void_p * DRV_MAX_LIGHTS),
        ("specular_mips", C.c_void_p), ("specular_total_size", u32), ("specular_levels", u32),
    ]


class HostFrame(C.Structure):
    """``drv_host_frame`` (drv_gi.h)."""
    _fields_ = [
        ("depth", C.c_void_p), ("normal_rg16i", C.c_void_p), ("diffuse_srgb8x", C.c_void_p),
        ("num_lights", u32),
        ("rsm_flux_rgbx16f", C.c_void_p * DRV_MAX_LIGHTS), ("rsm_normal_rg16i", C.c_void_p * DRV_MAX_LIGHTS),
        ("rsm_depthlinsq_rg16f", C.c_void_p * DRV_MAX_LIGHTS), ("rsm_resolution", u32 * DRV_MAX_LIGHTS),
        ("hdr_out", C.c_void_p), ("bands", u32),
    ]


assert C.sizeof(Constant) == 80
assert C.sizeof(PerFrame) == 288
assert C.sizeof(CAVCascade) == 64
assert C.sizeof(VolumeInfo) == 288
assert C.sizeof(SpotLight) == 224
assert C.sizeof(CacheCounter) == 16
assert SpotLight.LightPosition.offset == 32 and SpotLight.RSMRenderResolution.offset == 192
assert PerFrame.CameraPosition.offset == 256 and PerFrame.PassedTime.offset == 284

# numpy views of the buffer element types
VPL_DTYPE = np.dtype([("Position", "<f4", 3), ("DiscArea", "<f4"), ("Normal", "<f4", 3), ("_pad0", "<f4"),
                      ("Flux", "<f4", 3), ("_pad1", "<f4")])
SHADOW_BLOCK_DTYPE = np.dtype([("AverageValPos", "<f4", 3), ("DistToSphereRad", "<f4")])
assert VPL_DTYPE.itemsize == 48 and SHADOW_BLOCK_DTYPE.itemsize == 16


def entry_stride(sh_order: int) -> int:
    """LightCacheEntry size: 64 B (SH1) / 128 B (SH2), lightcache.glsl:33-57."""
    return 64 if sh_order == 1 else 128


def voxel_levels(res: int) -> int:
    """floor(log2(res)) + 1 mip levels (glhelper/texture.cpp:54-71)."""
    return int(res).bit_length()


def voxel_level_offset(res: int, level: int) -> int:
    off = 0
    for _ in range(level):
        off += res ** 3
        res //= 2
    return off


def voxel_chain_bytes(res: int) -> int:
    return voxel_level_offset(res, voxel_levels(res))


def rsm_level_offset(res: int, level: int) -> int:
    """Texel offset of mip ``level`` (>= 1) in a context-owned RSM mip buffer."""
    off = 0
    for l in range(1, level):
        off += (res >> l) ** 2
    return off
