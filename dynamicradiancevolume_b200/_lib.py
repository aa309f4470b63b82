"""ctypes binding of ``libdrv_gi.so`` (the C-ABI declared in ``include/drv_gi.h``).

The library is built in-tree by ``dynamicradiancevolume_b200.build``; there is
no Python or CPU fallback: if the shared object is missing, or no CUDA device is
present when a context is created, the call fails loudly.
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrv_gi.so")

# every symbol include/drv_gi.h declares: (name, restype, argtypes)
_P = C.c_void_p
_u32 = C.c_uint32
_i32 = C.c_int32
_f32 = C.c_float
_st = C.c_int
SYMBOLS = [
    ("drv_create", _st, [C.POINTER(abi.Config), C.POINTER(_P)]),
    ("drv_destroy", None, [_P]),
    ("drv_last_error", C.c_char_p, [_P]),
    ("drv_version", C.c_char_p, []),
    ("drv_set_constant", _st, [_P, C.POINTER(abi.Constant)]),
    ("drv_set_per_frame", _st, [_P, C.POINTER(abi.PerFrame)]),
    ("drv_set_volume_info", _st, [_P, C.POINTER(abi.VolumeInfo)]),
    ("drv_set_light_count", _st, [_P, _u32]),
    ("drv_set_spot_light", _st, [_P, _u32, C.POINTER(abi.SpotLight)]),
    ("drv_bind_gbuffer", _st, [_P, _P, _P, _P, _u32, _u32]),
    ("drv_bind_gbuffer_material", _st, [_P, _P]),
    ("drv_prepare_specular_envmaps", _st, [_P]),
    ("drv_bind_rsm", _st, [_P, _u32, _P, _P, _P, _u32]),
    ("drv_prepare_rsm", _st, [_P, _u32]),
    ("drv_voxelize", _st, [_P, _P, _u32, C.POINTER(_f32 * 16), _f32, _u32]),
    ("drv_set_voxel_volume", _st, [_P, _P]),
    ("drv_allocate_caches", _st, [_P]),
    ("drv_light_caches", _st, [_P]),
    ("drv_apply_caches", _st, [_P, _P, _u32]),
    ("drv_apply_caches_rows", _st, [_P, _P, _u32, _u32, _u32]),
    ("drv_draw", _st, [_P, _P, _u32]),
    ("drv_draw_frame", _st, [_P, _P, _u32, _u32]),
    ("drv_graph_stats", _st, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("drv_live_vpl_counts", _st, [_P, _P]),
    ("drv_fill_rsm", _st, [_P, _u32, _P, _P, _P, _P, _u32]),
    ("drv_cone_trace_ao", _st, [_P, _P]),
    ("drv_tonemap", _st, [_P, _P, C.c_float, C.c_float, _P]),
    ("drv_write_pfm", _st, [C.c_char_p, _P, _u32, _u32]),
    ("drv_save_to_pfm", _st, [_P, _P, C.c_char_p]),
    ("drv_bind_scene", _st, [_P, _P, _u32, _P, C.c_float]),
    ("drv_export_hdr_ipc", _st, [_P, _P]),
    ("drv_import_peer_hdr", _st, [_P, _u32, _P]),
    ("drv_get_buffers", _st, [_P, C.POINTER(abi.Buffers)]),
    ("drv_rsm_level_offset", C.c_uint64, [_u32, _u32]),
    ("drv_voxel_level_offset", C.c_uint64, [_u32, _u32]),
    ("drv_active_cache_count", _st, [_P, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    ("drv_set_synthetic_entries", _st, [_P, _P, _u32]),
    ("drv_set_vpls", _st, [_P, _u32, _P, _u32]),
    ("drv_set_shard", _st, [_P, _u32, _u32]),
    ("drv_shard_range", None, [_u32, _u32, _u32, C.POINTER(_u32), C.POINTER(_u32)]),
    ("drv_set_shard_interleave", _st, [_P, _u32]),
    ("drv_shard_entry", None, [_u32, _u32, _u32, C.POINTER(_u32)]),
    ("drv_export_entries_ipc", _st, [_P, C.POINTER(C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)]),
    ("drv_import_peer_entries", _st, [_P, _u32, C.POINTER(C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)]),
    ("drv_peer_barrier", _st, [_P]),
    ("drv_peer_status", _st, [_P, C.POINTER(_u32), C.POINTER(_u32)]),
    ("drv_peer_reset", _st, [_P]),
    ("drv_enable_stage_timers", _st, [_P, C.c_int]),
    ("drv_stage_ms", _st, [_P, C.c_int, C.POINTER(_f32)]),
    ("drv_stage_name", C.c_char_p, [C.c_int]),
    ("drv_kernel_launches", C.c_uint64, [_P]),
    ("drv_upload_gbuffer", _st, [_P, _P, _P, _P, _u32, _u32]),
    ("drv_upload_rsm", _st, [_P, _u32, _P, _P, _P, _u32]),
    ("drv_draw_to_host", _st, [_P, _P]),
    ("drv_draw_host_frame", _st, [_P, C.POINTER(abi.HostFrame)]),
    ("drv_pack_constant", None, [C.POINTER(abi.Constant), _i32, _i32, _i32, _i32, _i32, _u32]),
    ("drv_pack_specular", None, [C.POINTER(abi.Constant), _u32, _u32]),
    ("drv_pack_per_frame", None, [C.POINTER(abi.PerFrame), _P, _f32]),
    ("drv_pack_volume_info", None, [C.POINTER(abi.VolumeInfo), _P, C.POINTER(_f32 * 3), C.POINTER(_f32 * 3), _i32, _i32,
                                     _i32, C.POINTER(_f32), _f32]),
    ("drv_pack_spot_light", None, [C.POINTER(abi.SpotLight), _P]),
    ("drv_microbench", _st, [_i32, _u32, C.POINTER(C.c_double)]),
    ("drv_microbench_name", C.c_char_p, [_u32]),
    ("drv_microbench_count", _u32, []),
    ("drv_debug_gather_trace", _st, [_P, _P, _u32, C.POINTER(_u32)]),
    ("drv_debug_cone_steps", _st, [_P, C.POINTER(C.c_uint64)]),
    ("drv_debug_host_frame_timeline", _st, [_P, C.POINTER(_f32), _u32, C.POINTER(_u32)]),
]

# the pure-host part of the C-ABI (uniform-block packers): also built as libdrv_host.so with plain g++, so that host
# tooling which only prepares inputs — the CPU reference arm of bench.py, the oracle tests — never maps the CUDA library
HOST_LIB_PATH = os.path.join(_HERE, "libdrv_host.so")
HOST_SYMBOLS = [s for s in SYMBOLS if s[0].startswith("drv_pack_")]

_lib = None
_host = None


class DrvError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libdrv_gi error %d: %s" % (status, message))
        self.status = status


def load():
    """Load libdrv_gi.so (once) and attach prototypes. Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdrv_gi.so is missing (%s). Build it with `python -m dynamicradiancevolume_b200.build` "
            "or __graft_entry__.build(); there is no CPU fallback for the CUDA path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_host():
    """Load libdrv_host.so (once): the drv_pack_* entry points compiled without CUDA."""
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError("libdrv_host.so is missing (%s). Build it with `python -m dynamicradiancevolume_b200.build`."
                          % HOST_LIB_PATH)
    lib = C.CDLL(HOST_LIB_PATH)
    for name, res, args in HOST_SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _host = lib
    return lib


class CameraDesc(C.Structure):
    """``drv_camera_desc`` (camera/camera.hpp:15-39)."""
    _fields_ = [("position", _f32 * 3), ("direction", _f32 * 3), ("up", _f32 * 3), ("hfov_degrees", _f32),
                ("aspect_ratio", _f32), ("near_plane", _f32), ("far_plane", _f32)]


class LightDesc(C.Structure):
    """``drv_light_desc`` (scene/light.hpp:8-55)."""
    _fields_ = [("intensity", _f32 * 3), ("position", _f32 * 3), ("direction", _f32 * 3), ("half_angle", _f32),
                ("rsm_resolution", _u32), ("rsm_read_lod", _u32), ("normal_offset_shadow_bias", _f32),
                ("shadow_bias", _f32), ("indirect_shadow_lod", _u32), ("near_plane", _f32), ("far_plane", _f32)]
