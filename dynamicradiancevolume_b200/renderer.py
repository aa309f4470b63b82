"""Host-side mirror of the reference's ``class Renderer`` for the indirect-lighting path.

Method names, argument meaning and defaults follow
``DynamicRadianceVolume/rendering/renderer.hpp:36-216`` (``SetMaxCacheCount``,
``SetCAVCascades``, ``SetIndirectDiffuseMode``, ``AllocateCaches``,
``LightCachesRSM``, ``ApplyCaches``, ``Draw`` ...), so a test or a caller written
against the reference reads the same here. Everything below the method bodies is
the C-ABI of ``libdrv_gi`` (``include/drv_gi.h``); PyTorch only provides device
memory and the CUDA stream. There is no CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib, abi


@dataclass
class Camera:
    """``Camera`` (camera/camera.hpp:15-39); defaults of application.cpp:51-52."""
    position: Sequence[float] = (0.0, 2.5, 5.0)
    direction: Sequence[float] = (0.0, -2.5, -5.0)
    up: Sequence[float] = (0.0, 1.0, 0.0)
    hfov_degrees: float = 60.0
    aspect_ratio: float = 16.0 / 9.0
    near_plane: float = 0.1
    far_plane: float = 1000.0

    def desc(self) -> "_lib.CameraDesc":
        d = _lib.CameraDesc()
        d.position[:] = self.position
        d.direction[:] = self.direction
        d.up[:] = self.up
        d.hfov_degrees = self.hfov_degrees
        d.aspect_ratio = self.aspect_ratio
        d.near_plane = self.near_plane
        d.far_plane = self.far_plane
        return d


@dataclass
class Light:
    """``struct Light`` (scene/light.hpp:8-55) with its constructor defaults."""
    intensity: Sequence[float] = (10.0, 10.0, 10.0)
    position: Sequence[float] = (0.0, 0.0, 0.0)
    direction: Sequence[float] = (0.0, 0.0, 1.0)
    halfAngle: float = 0.5
    rsmResolution: int = 1024
    rsmReadLod: int = 4
    normalOffsetShadowBias: float = 0.01
    shadowBias: float = 0.0001
    indirectShadowComputationLod: int = 2
    nearPlane: float = 0.1      # Light::nearPlane, scene/scene.cpp:6
    farPlane: float = 10000.0   # Light::farPlane, scene/scene.cpp:7

    def desc(self) -> "_lib.LightDesc":
        d = _lib.LightDesc()
        d.intensity[:] = self.intensity
        d.position[:] = self.position
        d.direction[:] = self.direction
        d.half_angle = self.halfAngle
        d.rsm_resolution = self.rsmResolution
        d.rsm_read_lod = self.rsmReadLod
        d.normal_offset_shadow_bias = self.normalOffsetShadowBias
        d.shadow_bias = self.shadowBias
        d.indirect_shadow_lod = self.indirectShadowComputationLod
        d.near_plane = self.nearPlane
        d.far_plane = self.farPlane
        return d


@dataclass
class Scene:
    """The slice of ``class Scene`` (scene/scene.hpp:28-41) the path consumes: lights, the
    bounding box, and the entity triangles (world matrix + positions) for the voxeliser."""
    lights: List[Light] = field(default_factory=list)
    bbox_min: Sequence[float] = (0.0, 0.0, 0.0)
    bbox_max: Sequence[float] = (1.0, 1.0, 1.0)
    entities: list = field(default_factory=list)  # [(device float tensor [nTris*9], world 4x4 row-major)]


class IndirectDiffuseMode:
    SH1 = 1
    SH2 = 2


# ---- uniform-block packers (pure host; usable without a GPU) -------------------------------

def pack_constant(width, height, voxel_res, cav_res, cav_cascades, max_caches) -> abi.Constant:
    """≙ Renderer::UpdateConstantUBO (renderer.cpp:290-322)."""
    out = abi.Constant()
    _lib.load_host().drv_pack_constant(C.byref(out), width, height, voxel_res, cav_res, cav_cascades, max_caches)
    return out


def pack_specular(constant: abi.Constant, max_caches: int, per_cache_size: int = 16) -> abi.Constant:
    """Fills the four specular environment-map fields of a Constant block (renderer.cpp:253, 316-319)."""
    _lib.load_host().drv_pack_specular(C.byref(constant), max_caches, per_cache_size)
    return constant


def pack_per_frame(camera: Camera, passed_time: float = 0.0) -> abi.PerFrame:
    """≙ Renderer::UpdatePerFrameUBO (renderer.cpp:324-344)."""
    out = abi.PerFrame()
    d = camera.desc()
    _lib.load_host().drv_pack_per_frame(C.byref(out), C.addressof(d), passed_time)
    return out


def pack_volume_info(camera: Camera, bbox_min, bbox_max, voxel_res, cav_res, cascade_world_sizes,
                     transition_zone_size) -> abi.VolumeInfo:
    """≙ Renderer::UpdateVolumeUBO (renderer.cpp:346-431)."""
    out = abi.VolumeInfo()
    d = camera.desc()
    mn = (C.c_float * 3)(*bbox_min)
    mx = (C.c_float * 3)(*bbox_max)
    sizes = (C.c_float * len(cascade_world_sizes))(*cascade_world_sizes)
    _lib.load_host().drv_pack_volume_info(C.byref(out), C.addressof(d), C.byref(mn), C.byref(mx), voxel_res, cav_res,
                                     len(cascade_world_sizes), sizes, transition_zone_size)
    return out


def pack_spot_light(light: Light) -> abi.SpotLight:
    """≙ Renderer::PrepareLights (renderer.cpp:664-725)."""
    out = abi.SpotLight()
    d = light.desc()
    _lib.load_host().drv_pack_spot_light(C.byref(out), C.addressof(d))
    return out


def default_cascade_world_sizes(num_cascades: int, first: float = 4.0) -> List[float]:
    """Renderer::SetCAVCascades defaults: 4, 8, 16, ... (renderer.cpp:1181-1187)."""
    return [first * (2.0 ** i) for i in range(num_cascades)]


def shard_range(count: int, rank: int, world: int):
    b, e = C.c_uint32(), C.c_uint32()
    _lib.load().drv_shard_range(count, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


def shard_count(count: int, rank: int, world: int, interleave: bool = False) -> int:
    """Entries rank `rank` lights: the contiguous range of ``shard_range`` or, interleaved, every world-th 64-entry
    group of the cell-ordered list (``drv_set_shard_interleave``)."""
    if not interleave or world <= 1:
        b, e = shard_range(count, rank, world)
        return e - b
    groups = (count + 63) // 64
    owned = (groups - rank + world - 1) // world if groups > rank else 0
    n = owned * 64
    if owned and (groups - 1) % world == rank:
        n -= groups * 64 - count
    return n



class Context:
    """Thin RAII wrapper of ``drv_ctx`` — one per (configuration, device)."""

    def __init__(self, *, max_cache_count=16384, cav_cascades=3, cav_resolution=32, voxel_resolution=128, sh_order=1,
                 indirect_shadow=True, cascade_transitions=True, width=1920, height=1080, max_lights=1,
                 max_rsm_resolution=1024, device=0, stream=None, gather_variant=0, indirect_specular=False,
                 specular_per_cache_size=16, specular_fill_holes_level=0):
        self.lib = _lib.load()
        cfg = abi.Config()
        cfg.indirect_specular = 1 if indirect_specular else 0
        cfg.specular_per_cache_size = specular_per_cache_size
        cfg.specular_fill_holes_level = specular_fill_holes_level
        cfg.max_cache_count = max_cache_count
        cfg.cav_cascades = cav_cascades
        cfg.cav_resolution = cav_resolution
        cfg.voxel_resolution = voxel_resolution
        cfg.sh_order = sh_order
        cfg.indirect_shadow = 1 if indirect_shadow else 0
        cfg.cascade_transitions = 1 if cascade_transitions else 0
        cfg.backbuffer_width = width
        cfg.backbuffer_height = height
        cfg.max_lights = max_lights
        cfg.max_rsm_resolution = max_rsm_resolution
        cfg.device = device
        cfg.stream = stream
        cfg.gather_variant = gather_variant
        self.cfg = cfg
        self.handle = C.c_void_p()
        st = self.lib.drv_create(C.byref(cfg), C.byref(self.handle))
        if st != abi.DRV_OK:
            self.handle = C.c_void_p()
            raise _lib.DrvError(st, (self.lib.drv_last_error(None) or b"").decode())
        self.entry_stride = abi.entry_stride(sh_order)
        self._keep = []  # tensors whose storage the context borrows

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.drv_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        if st != abi.DRV_OK:
            raise _lib.DrvError(st, (self.lib.drv_last_error(self.handle) or b"").decode())

    # -- thin 1:1 calls --
    def set_constant(self, b): self.check(self.lib.drv_set_constant(self.handle, C.byref(b)))
    def set_per_frame(self, b): self.check(self.lib.drv_set_per_frame(self.handle, C.byref(b)))
    def set_volume_info(self, b): self.check(self.lib.drv_set_volume_info(self.handle, C.byref(b)))
    def set_light_count(self, n): self.check(self.lib.drv_set_light_count(self.handle, n))
    def set_spot_light(self, i, b): self.check(self.lib.drv_set_spot_light(self.handle, i, C.byref(b)))

    def bind_gbuffer_material(self, rough_metal):
        """RG8 roughness / metallic plane (device tensor [H, W, 2] uint8): read only with indirect specular."""
        self._keep_rm = rough_metal
        self.check(self.lib.drv_bind_gbuffer_material(self.handle, rough_metal.data_ptr()))

    def prepare_specular_envmaps(self): self.check(self.lib.drv_prepare_specular_envmaps(self.handle))

    def read_specular_mips(self):
        """Every level of the environment-map atlas (uint32 R11F_G11F_B10F texels), level 0 first."""
        b = self.buffers()
        n = sum((b.specular_total_size >> l) ** 2 for l in range(b.specular_levels))
        return self._read(b.specular_mips, n * 4).view(np.uint32).copy()

    def bind_gbuffer(self, depth, normal, diffuse):
        h, w = depth.shape[0], depth.shape[1]
        self._keep_gb = (depth, normal, diffuse)
        self.check(self.lib.drv_bind_gbuffer(self.handle, depth.data_ptr(), normal.data_ptr(), diffuse.data_ptr(), w, h))

    def bind_rsm(self, light, flux, normal, depth):
        res = flux.shape[0]
        self._keep.append((flux, normal, depth))
        self.check(self.lib.drv_bind_rsm(self.handle, light, flux.data_ptr(), normal.data_ptr(), depth.data_ptr(), res))

    def prepare_rsm(self, light): self.check(self.lib.drv_prepare_rsm(self.handle, light))

    def voxelize(self, tris, world=None, adaption=1.0, flags=abi.DRV_VOXELIZE_CLEAR | abi.DRV_VOXELIZE_FINISH):
        w = (C.c_float * 16)(*(world if world is not None else np.eye(4, dtype=np.float32).ravel().tolist()))
        n = 0 if tris is None else tris.numel() // 9
        ptr = None if tris is None else tris.data_ptr()
        self.check(self.lib.drv_voxelize(self.handle, ptr, n, C.byref(w), adaption, flags))

    def set_voxel_volume(self, level0):
        """level0: uint8 tensor / array of voxel_resolution^3 bytes (host or device)."""
        ptr = level0.data_ptr() if hasattr(level0, "data_ptr") else level0.ctypes.data
        self._keep_vol = level0
        self.check(self.lib.drv_set_voxel_volume(self.handle, ptr))

    def allocate_caches(self): self.check(self.lib.drv_allocate_caches(self.handle))
    def light_caches(self): self.check(self.lib.drv_light_caches(self.handle))
    def apply_caches(self, out, fmt): self.check(self.lib.drv_apply_caches(self.handle, out.data_ptr(), fmt))
    def draw(self, out, fmt): self.check(self.lib.drv_draw(self.handle, out.data_ptr(), fmt))

    def bind_scene(self, tris, world=None, adaption=1.0):
        """Geometry for ``DRV_FRAME_VOXELIZE`` (``drv_bind_scene``); ``tris`` is a device tensor of n*9 floats."""
        w = (C.c_float * 16)(*(world if world is not None else np.eye(4, dtype=np.float32).ravel().tolist()))
        n = 0 if tris is None else tris.numel() // 9
        self._keep_scene = tris
        self.check(self.lib.drv_bind_scene(self.handle, None if tris is None else tris.data_ptr(), n, C.byref(w), adaption))

    # -- rows next to the hot path --
    def fill_rsm(self, light, position, normal, basecolor, coverage=None):
        """``drv_fill_rsm``: device tensors [R, R, 3] float32 (+ optional [R, R] uint8 coverage); binds the result."""
        res = position.shape[0]
        self._keep_fill = (position, normal, basecolor, coverage)
        self.check(self.lib.drv_fill_rsm(self.handle, light, position.data_ptr(), normal.data_ptr(), basecolor.data_ptr(),
                                         None if coverage is None else coverage.data_ptr(), res))

    def cone_trace_ao(self, out):
        """``drv_cone_trace_ao`` into a device tensor [H, W] float32 (discarded pixels are left untouched)."""
        self.check(self.lib.drv_cone_trace_ao(self.handle, out.data_ptr()))

    def tonemap(self, hdr16, out32, exposure=1.0, l_max=1.2):
        """``drv_tonemap``: RGBA16F [H, W, 4] -> float32 [H, W, 4]."""
        self.check(self.lib.drv_tonemap(self.handle, hdr16.data_ptr(), exposure, l_max, out32.data_ptr()))

    def save_to_pfm(self, hdr16, path):
        self.check(self.lib.drv_save_to_pfm(self.handle, hdr16.data_ptr(), path.encode()))

    def live_vpl_counts(self):
        """VPLs with non-zero flux per light — what the gather streams (``drv_live_vpl_counts``)."""
        a = (C.c_uint32 * 16)()
        self.check(self.lib.drv_live_vpl_counts(self.handle, a))
        return list(a)

    def draw_frame(self, out, fmt, flags=abi.DRV_FRAME_PREPARE_RSM):
        """Whole frame in GPU order (``drv_draw_frame``): light side || allocation, join, gather, apply."""
        self.check(self.lib.drv_draw_frame(self.handle, None if out is None else out.data_ptr(), fmt, flags))

    def apply_caches_rows(self, out, fmt, y0, y1):
        self.check(self.lib.drv_apply_caches_rows(self.handle, out.data_ptr(), fmt, y0, y1))
    def set_shard(self, rank, world): self.check(self.lib.drv_set_shard(self.handle, rank, world))
    def set_shard_interleave(self, on=True): self.check(self.lib.drv_set_shard_interleave(self.handle, 1 if on else 0))
    def enable_stage_timers(self, on=True): self.check(self.lib.drv_enable_stage_timers(self.handle, 1 if on else 0))
    def kernel_launches(self): return int(self.lib.drv_kernel_launches(self.handle))

    def cone_steps(self):
        """Voxel samples the cone pass has taken since the last call (resets the count)."""
        n = C.c_uint64()
        self.check(self.lib.drv_debug_cone_steps(self.handle, C.byref(n)))
        return n.value

    def host_frame_timeline(self):
        """Timeline of the last draw_host_frame, ms after its first copy was queued: {"rsm_in", "depth_in", "lit",
        "bands": [(inputs_in, applied, copied_out), ...]}."""
        buf = (C.c_float * (3 + 3 * 32))()
        n = C.c_uint32()
        self.check(self.lib.drv_debug_host_frame_timeline(self.handle, buf, len(buf), C.byref(n)))
        return {"rsm_in": buf[0], "depth_in": buf[1], "lit": buf[2],
                "bands": [(buf[3 + 3 * b], buf[4 + 3 * b], buf[5 + 3 * b]) for b in range(n.value)]}

    def graph_stats(self):
        """(instantiations, in-place updates) of the frame graph of draw_frame(DRV_FRAME_GRAPH)."""
        a, b = C.c_uint64(), C.c_uint64()
        self.check(self.lib.drv_graph_stats(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def peer_status(self):
        """Raises DrvError(DRV_ERR_PEER) if a cross-GPU barrier of an earlier frame timed out."""
        e, r = C.c_uint32(), C.c_uint32()
        self.check(self.lib.drv_peer_status(self.handle, C.byref(e), C.byref(r)))

    def peer_reset(self): self.check(self.lib.drv_peer_reset(self.handle))

    def gather_trace(self):
        """Diagnostics (gather_variant bit 18): [ctas, 8] uint64: 4 %globaltimer stamps, %smid, 3 clock64 values of the last pair-kernel launch."""
        import numpy as np
        out = np.zeros((8192, 8), np.uint64)
        n = C.c_uint32()
        self.check(self.lib.drv_debug_gather_trace(self.handle, out.ctypes.data, 8192, C.byref(n)))
        return out[:n.value]

    def stage_ms(self, stage: int) -> float:
        ms = C.c_float()
        self.check(self.lib.drv_stage_ms(self.handle, stage, C.byref(ms)))
        return ms.value

    def buffers(self) -> abi.Buffers:
        b = abi.Buffers()
        self.check(self.lib.drv_get_buffers(self.handle, C.byref(b)))
        return b

    def active_cache_count(self):
        n, ov, oob = C.c_uint32(), C.c_uint32(), C.c_uint32()
        st = self.lib.drv_active_cache_count(self.handle, C.byref(n), C.byref(ov), C.byref(oob))
        if st not in (abi.DRV_OK, abi.DRV_ERR_CAPACITY):
            self.check(st)
        return n.value, ov.value, oob.value

    def set_synthetic_entries(self, positions):
        self._keep_syn = positions
        self.check(self.lib.drv_set_synthetic_entries(self.handle, positions.data_ptr(), positions.shape[0]))

    def set_vpls(self, light, vpls_ptr, n): self.check(self.lib.drv_set_vpls(self.handle, light, vpls_ptr, n))

    def upload_gbuffer(self, depth, normal, diffuse):
        h, w = depth.shape[0], depth.shape[1]
        self.check(self.lib.drv_upload_gbuffer(self.handle, depth.data_ptr(), normal.data_ptr(), diffuse.data_ptr(), w, h))

    def upload_rsm(self, light, flux, normal, depth):
        self.check(self.lib.drv_upload_rsm(self.handle, light, flux.data_ptr(), normal.data_ptr(), depth.data_ptr(),
                                           flux.shape[0]))

    def draw_to_host(self, hdr_host): self.check(self.lib.drv_draw_to_host(self.handle, hdr_host.data_ptr()))

    def draw_host_frame(self, depth, normal, diffuse, rsms, hdr_host, bands=0):
        """Pipelined end-to-end frame from pinned host tensors (``drv_draw_host_frame``). rsms: [(flux, normal, depthLinSq)]."""
        f = abi.HostFrame()
        f.depth, f.normal_rg16i, f.diffuse_srgb8x = depth.data_ptr(), normal.data_ptr(), diffuse.data_ptr()
        f.num_lights = len(rsms)
        for i, (fl, n, d) in enumerate(rsms):
            f.rsm_flux_rgbx16f[i], f.rsm_normal_rg16i[i], f.rsm_depthlinsq_rg16f[i] = fl.data_ptr(), n.data_ptr(), d.data_ptr()
            f.rsm_resolution[i] = fl.shape[0]
        f.hdr_out = hdr_host.data_ptr()
        f.bands = bands
        self.check(self.lib.drv_draw_host_frame(self.handle, C.byref(f)))

    def peer_barrier(self): self.check(self.lib.drv_peer_barrier(self.handle))

    def export_entries_ipc(self) -> bytes:
        h = (C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)()
        self.check(self.lib.drv_export_entries_ipc(self.handle, C.byref(h)))
        return bytes(h)

    def import_peer_entries(self, rank: int, handle: bytes):
        h = (C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)(*handle)
        self.check(self.lib.drv_import_peer_entries(self.handle, rank, C.byref(h)))

    def export_hdr_ipc(self) -> bytes:
        h = (C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)()
        self.check(self.lib.drv_export_hdr_ipc(self.handle, C.byref(h)))
        return bytes(h)

    def import_peer_hdr(self, rank: int, handle: bytes):
        h = (C.c_uint8 * abi.DRV_IPC_HANDLE_BYTES)(*handle)
        self.check(self.lib.drv_import_peer_hdr(self.handle, rank, C.byref(h)))

    def hdr16_tensor(self):
        """The context-owned RGBA16F target as a torch tensor [H, W, 4] (rank 0's holds the gathered image)."""
        import torch
        b = self.buffers()
        if not b.hdr16:
            raise DrvError(abi.DRV_ERR_NOT_BOUND, "no context-owned HDR target yet")
        w, h = self.cfg.backbuffer_width, self.cfg.backbuffer_height
        return self.device_view(b.hdr16, w * h * 8).view(torch.float16).view(h, w, 4)

    # -- device -> host readback helpers (parity tests) --
    def device_view(self, ptr, nbytes):
        """A torch uint8 tensor aliasing ``nbytes`` of context-owned device memory at ``ptr``."""
        import torch

        class _Span:
            pass
        span = _Span()
        span.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}
        return torch.as_tensor(span, device="cuda:%d" % self.cfg.device)

    def entries_tensor(self):
        """The whole LightCacheBuffer as a [max_cache_count, stride/4] float32 device tensor (for collectives)."""
        import torch
        b = self.buffers()
        return self.device_view(b.entries, b.max_cache_count * b.entry_stride).view(torch.float32).view(
            b.max_cache_count, b.entry_stride // 4)

    def _read(self, ptr, nbytes):
        import torch
        torch.cuda.synchronize(self.cfg.device)
        return self.device_view(ptr, nbytes).cpu().numpy()

    def read_entries(self, count=None) -> np.ndarray:
        b = self.buffers()
        n = self.active_cache_count()[0] if count is None else count
        raw = self._read(b.entries, n * b.entry_stride)
        return raw.view(np.float32).reshape(n, b.entry_stride // 4).copy()

    def read_atlas(self) -> np.ndarray:
        b = self.buffers()
        raw = self._read(b.cav_atlas, b.cav_width * b.cav_height * b.cav_depth * 4)
        return raw.view(np.uint32).reshape(b.cav_depth, b.cav_height, b.cav_width).copy()

    def read_vpls(self, light, n) -> np.ndarray:
        b = self.buffers()
        return self._read(b.vpls[light], n * 48).view(abi.VPL_DTYPE).copy()

    def read_shadow_blocks(self, light, n) -> np.ndarray:
        b = self.buffers()
        return self._read(b.shadow_blocks[light], n * 16).view(abi.SHADOW_BLOCK_DTYPE).copy()

    def read_voxel_chain(self) -> np.ndarray:
        b = self.buffers()
        return self._read(b.voxel_chain, int(b.voxel_chain_bytes)).copy()

    def read_voxel_target(self) -> np.ndarray:
        b = self.buffers()
        return self._read(b.voxel_target, b.voxel_resolution ** 3).copy()

    def read_rsm_mip(self, light, res, level):
        """(flux[h,h,4] u16, normal[h,h,2] i16, depth[h,h,2] u16) of context-owned mip ``level`` >= 1."""
        b = self.buffers()
        off = abi.rsm_level_offset(res, level)
        h = res >> level
        f = self._read(b.rsm_flux_mips[light] + off * 8, h * h * 8).view(np.uint16).reshape(h, h, 4).copy()
        n = self._read(b.rsm_normal_mips[light] + off * 4, h * h * 4).view(np.int16).reshape(h, h, 2).copy()
        d = self._read(b.rsm_depth_mips[light] + off * 4, h * h * 4).view(np.uint16).reshape(h, h, 2).copy()
        return f, n, d


class Renderer:
    """The reference's ``Renderer`` interface for the DYN_RADIANCE_VOLUME path."""

    s_maxNumCAVCascades = abi.DRV_MAX_CASCADES  # renderer.hpp:146

    def __init__(self, scene: Scene, resolution, device: int = 0, stream=None, gather_variant: int = 0):
        # constructor defaults, renderer.cpp:36-51, 86-90
        self.m_scene = scene
        self.m_indirectDiffuseMode = IndirectDiffuseMode.SH1
        self.m_CAVCascadeTransitionSize = 2.0
        self.m_indirectShadow = True
        self.m_maxNumLightCaches = 16384
        self.m_CAVCascadeWorldSize: List[float] = []
        self.m_cavResolution = 0
        self.m_voxelResolution = 128
        self.m_adaptionRate = 10.0  # Voxelization ctor, voxelization.cpp:25
        self.m_readLightCacheCount = False
        self.m_lastNumLightCaches = 0
        self.m_passedTime = 0.0
        self.m_resolution = tuple(resolution)
        self._device = device
        # All stage kernels run on ONE stream (the reference issues everything on the GL context's queue);
        # tensor helpers below (HDR clear) are issued on the same stream so no cross-stream ordering is needed.
        self._tstream = None
        if stream is None:
            import torch
            self._tstream = torch.cuda.Stream(device=device)
            stream = self._tstream.cuda_stream
        self._stream = stream
        self._variant = gather_variant
        self._ctx: Optional[Context] = None
        self._gbuffer = None
        self._rsms = {}
        self._hdr = None
        self._voxel_pending = 0.0
        # indirect specular (renderer.cpp:43-45, 49): off by default, 16^2 texels per cache, no hole filling
        self.m_indirectSpecular = False
        self.m_specularEnvmapPerCacheSize = 16
        self.m_specularEnvmapMaxFillHolesLevel = 0
        self.m_specularEnvmapDirectWrite = True
        self._rough_metal = None
        self.SetCAVCascades(3, 32)

    # ---- setters (renderer.hpp:63-141): each invalidates the context like a shader reload ----
    def _invalidate(self):
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = None

    def SetIndirectDiffuseMode(self, mode): self.m_indirectDiffuseMode = mode; self._invalidate()
    def GetIndirectDiffuseMode(self): return self.m_indirectDiffuseMode
    def SetIndirectShadow(self, active): self.m_indirectShadow = bool(active); self._invalidate()
    def GetIndirectShadow(self): return self.m_indirectShadow
    def SetVoxelVolumeResultion(self, resolution): self.m_voxelResolution = int(resolution); self._invalidate()
    def GetVoxelVolumeResultion(self): return self.m_voxelResolution
    def SetVoxelVolumeAdaptionRate(self, rate): self.m_adaptionRate = float(rate)
    def GetVoxelVolumeAdaptionRate(self): return self.m_adaptionRate
    def SetMaxCacheCount(self, n): self.m_maxNumLightCaches = int(n); self._invalidate()
    def GetMaxCacheCount(self): return self.m_maxNumLightCaches
    def OnScreenResize(self, resolution): self.m_resolution = tuple(resolution); self._invalidate()
    def SetScene(self, scene): self.m_scene = scene; self._invalidate()
    def GetScene(self): return self.m_scene
    def SetReadLightCacheCount(self, track): self.m_readLightCacheCount = bool(track); self.m_lastNumLightCaches = 0
    def GetReadLightCacheCount(self): return self.m_readLightCacheCount
    def GetLightCacheActiveCount(self): return self.m_lastNumLightCaches
    # indirect specular, renderer.hpp:80-106
    def SetIndirectSpecular(self, active): self.m_indirectSpecular = bool(active); self._invalidate()
    def GetIndirectSpecular(self): return self.m_indirectSpecular

    def SetPerCacheSpecularEnvMapSize(self, specularEnvmapPerCacheSize):
        """renderer.cpp:453-464: a power of two; the hole-fill level is clamped to log2(size)."""
        size = int(specularEnvmapPerCacheSize)
        assert size > 0 and size & (size - 1) == 0
        self.m_specularEnvmapPerCacheSize = size
        self.m_specularEnvmapMaxFillHolesLevel = min(self.m_specularEnvmapMaxFillHolesLevel, int(math.log2(size)))
        self._invalidate()

    def GetPerCacheSpecularEnvMapSize(self): return self.m_specularEnvmapPerCacheSize

    def SetSpecularEnvMapHoleFillLevel(self, holeFillLevel):
        self.m_specularEnvmapMaxFillHolesLevel = min(int(holeFillLevel), int(math.log2(self.m_specularEnvmapPerCacheSize)))
        self._invalidate()

    def GetSpecularEnvMapHoleFillLevel(self): return self.m_specularEnvmapMaxFillHolesLevel

    def SetSpecularEnvMapDirectWrite(self, directWrite):
        if not directWrite:  # the shared-exponent register path (#ifndef DIRECT_SPECULAR_MAP_WRITE) is out of scope, DESIGN 7
            raise NotImplementedError("only the reference's default, direct specular map write, is implemented")
        self.m_specularEnvmapDirectWrite = True

    def GetSpecularEnvMapDirectWrite(self): return self.m_specularEnvmapDirectWrite
    def GetCAVCascadeCount(self): return len(self.m_CAVCascadeWorldSize)
    def GetCAVResolution(self): return self.m_cavResolution
    def GetCAVCascadeTransitionSize(self): return self.m_CAVCascadeTransitionSize

    def GetCAVCascadeWorldSize(self, cascade):
        return float("nan") if cascade >= len(self.m_CAVCascadeWorldSize) else self.m_CAVCascadeWorldSize[cascade]

    def SetCAVCascades(self, numCascades, resolutionPerCascade):
        """renderer.cpp:1174-1193: keeps existing sizes, new cascades double the previous one."""
        assert 0 < numCascades <= self.s_maxNumCAVCascades and resolutionPerCascade > 0
        prev = list(self.m_CAVCascadeWorldSize)
        sizes = prev[:numCascades]
        if not sizes:
            sizes = [4.0]
        while len(sizes) < numCascades:
            sizes.append(sizes[-1] * 2.0)
        self.m_CAVCascadeWorldSize = sizes
        self.m_cavResolution = int(resolutionPerCascade)
        self._invalidate()

    def SetCAVCascadeWorldSize(self, cascade, size):
        assert cascade < len(self.m_CAVCascadeWorldSize) and size > 0.0
        self.m_CAVCascadeWorldSize[cascade] = float(size)

    def SetCAVCascadeTransitionSize(self, size):
        flip = (self.m_CAVCascadeTransitionSize > 0) != (size > 0)
        self.m_CAVCascadeTransitionSize = float(size)
        if flip:
            self._invalidate()

    # ---- inputs the reference rasterises itself (DrawSceneToGBuffer / DrawShadowMaps) ----
    def BindGBuffer(self, depth, normal, diffuse, roughnessMetallic=None):
        """Device tensors: depth f32 [H,W]; normal int16 [H,W,2]; diffuse uint8 [H,W,4]; roughnessMetallic uint8 [H,W,2]
        (the RG8 attachment, read only with indirect specular) (renderer.cpp:727-738)."""
        self._gbuffer = (depth, normal, diffuse)
        self._rough_metal = roughnessMetallic
        if self._ctx is not None:
            self._ctx.bind_gbuffer(depth, normal, diffuse)
            if roughnessMetallic is not None and self.m_indirectSpecular:
                self._ctx.bind_gbuffer_material(roughnessMetallic)

    def BindShadowMap(self, lightIndex, flux, normal, depthLinSq):
        """Level 0 of a light's RSM at rsmResolution (renderer.cpp:1288-1291)."""
        self._rsms[lightIndex] = (flux, normal, depthLinSq)
        if self._ctx is not None:
            self._ctx.bind_rsm(lightIndex, flux, normal, depthLinSq)
            self._ctx.prepare_rsm(lightIndex)

    # ---- context ----
    def context(self) -> Context:
        if self._ctx is None:
            lights = self.m_scene.lights
            max_rsm = max([1 << int(math.ceil(math.log2(l.rsmResolution))) for l in lights] + [16])
            self._ctx = Context(max_cache_count=self.m_maxNumLightCaches, cav_cascades=len(self.m_CAVCascadeWorldSize),
                                cav_resolution=self.m_cavResolution, voxel_resolution=self.m_voxelResolution,
                                sh_order=self.m_indirectDiffuseMode, indirect_shadow=self.m_indirectShadow,
                                cascade_transitions=self.m_CAVCascadeTransitionSize > 0.0, width=self.m_resolution[0],
                                height=self.m_resolution[1], max_lights=max(1, len(lights)), max_rsm_resolution=max_rsm,
                                device=self._device, stream=self._stream, gather_variant=self._variant,
                                indirect_specular=self.m_indirectSpecular,
                                specular_per_cache_size=self.m_specularEnvmapPerCacheSize,
                                specular_fill_holes_level=self.m_specularEnvmapMaxFillHolesLevel)
            if self._gbuffer is not None:
                self._ctx.bind_gbuffer(*self._gbuffer)
                if self._rough_metal is not None and self.m_indirectSpecular:
                    self._ctx.bind_gbuffer_material(self._rough_metal)
            for i, t in self._rsms.items():
                self._ctx.bind_rsm(i, *t)
                self._ctx.prepare_rsm(i)
            self.UpdateConstantUBO()
        return self._ctx

    # ---- packers ----
    def UpdateConstantUBO(self):
        c = pack_constant(self.m_resolution[0], self.m_resolution[1], self.m_voxelResolution, self.m_cavResolution,
                          len(self.m_CAVCascadeWorldSize), self.m_maxNumLightCaches)
        if self.m_indirectSpecular:  # SpecularEnvmap* members, renderer.cpp:317-319
            pack_specular(c, self.m_maxNumLightCaches, self.m_specularEnvmapPerCacheSize)
        self.m_constant = c
        if self._ctx is not None:
            self._ctx.set_constant(c)

    def UpdatePerFrameUBO(self, camera: Camera):
        self.m_perFrame = pack_per_frame(camera, self.m_passedTime)
        self.context().set_per_frame(self.m_perFrame)

    def UpdateVolumeUBO(self, camera: Camera):
        self.m_volumeInfo = pack_volume_info(camera, self.m_scene.bbox_min, self.m_scene.bbox_max, self.m_voxelResolution,
                                             self.m_cavResolution, self.m_CAVCascadeWorldSize,
                                             self.m_CAVCascadeTransitionSize)
        self.context().set_volume_info(self.m_volumeInfo)

    def PrepareLights(self):
        ctx = self.context()
        ctx.set_light_count(len(self.m_scene.lights))
        self.m_spotLights = []
        for i, l in enumerate(self.m_scene.lights):
            b = pack_spot_light(l)
            self.m_spotLights.append(b)
            ctx.set_spot_light(i, b)

    # ---- stages ----
    def VoxelizeScene(self, timeSinceLastBlend: float):
        """≙ Voxelization::VoxelizeScene (voxelization.cpp:90-176): adaption = floor(dt*rate*255)/255 with the
        remainder carried to the next frame."""
        self._voxel_pending += timeSinceLastBlend * self.m_adaptionRate * 255.0
        k = math.floor(self._voxel_pending)
        self._voxel_pending -= k
        if k <= 0:
            return
        adaption = min(k, 255) / 255.0
        ctx = self.context()
        ents = self.m_scene.entities
        if not ents:
            ctx.voxelize(None, None, adaption, abi.DRV_VOXELIZE_CLEAR | abi.DRV_VOXELIZE_FINISH)
        for i, (tris, world) in enumerate(ents):
            flags = (abi.DRV_VOXELIZE_CLEAR if i == 0 else 0) | (abi.DRV_VOXELIZE_FINISH if i == len(ents) - 1 else 0)
            ctx.voxelize(tris, world, adaption, flags)

    def AllocateCaches(self):
        ctx = self.context()
        if self.m_readLightCacheCount:  # one frame late, like renderer.cpp:960-966
            self.m_lastNumLightCaches = ctx.active_cache_count()[0]
        ctx.allocate_caches()

    def LightCachesRSM(self):
        self.context().light_caches()

    def PrepareSpecularEnvmaps(self):
        """Renderer::PrepareSpecularEnvmaps (renderer.cpp:994-1045): mip chain of the atlas, then hole filling."""
        self.context().prepare_specular_envmaps()

    def ApplyCaches(self, hdr=None, fmt=abi.DRV_HDR_RGBA16F_ADD):
        import torch
        if hdr is None:
            if self._hdr is None:
                with torch.cuda.stream(self._tstream if self._tstream is not None else torch.cuda.current_stream()):
                    self._hdr = torch.zeros(self.m_resolution[1], self.m_resolution[0], 4, dtype=torch.float16,
                                            device="cuda:%d" % self._device)
            hdr = self._hdr
        self.context().apply_caches(hdr, fmt)
        return hdr

    # -- the passes around the path: ConeTraceAO, tonemap, SaveToPFM --
    def SetExposure(self, exposure): self.m_tonemapExposure = float(exposure)        # renderer.cpp:1218-1221
    def GetExposure(self): return getattr(self, "m_tonemapExposure", 1.0)
    def SetTonemapLMax(self, tonemapLMax): self.m_tonemapLMax = float(tonemapLMax)   # renderer.cpp:1223-1227
    def GetTonemapLMax(self): return getattr(self, "m_tonemapLMax", 1.2)             # default renderer.cpp:39

    def ConeTraceAO(self, out=None):
        """Renderer::ConeTraceAO (renderer.cpp:936-949): [H, W] float32, discarded pixels keep 0."""
        import torch
        if out is None:
            out = torch.zeros(self.m_resolution[1], self.m_resolution[0], dtype=torch.float32, device="cuda:%d" % self._device)
        self.context().cone_trace_ao(out)
        return out

    def Tonemap(self, hdr=None, out=None):
        """The tonemap pass of Renderer::Draw (tonemapping.frag): RGBA16F HDR target -> float32 (rgb, 1)."""
        import torch
        hdr = self._hdr if hdr is None else hdr
        if out is None:
            out = torch.zeros(self.m_resolution[1], self.m_resolution[0], 4, dtype=torch.float32, device="cuda:%d" % self._device)
        self.context().tonemap(hdr, out, self.GetExposure(), self.GetTonemapLMax())
        return out

    def SaveToPFM(self, filename, hdr=None):
        """Renderer::SaveToPFM (renderer.cpp:1229-1235)."""
        self.context().save_to_pfm(self._hdr if hdr is None else hdr, filename)

    def Draw(self, camera: Camera, detachViewFromCameraUpdate: bool = False, timeSinceLastFrame: float = 0.0, hdr=None):
        """The DYN_RADIANCE_VOLUME case of Renderer::Draw (renderer.cpp:501-594) minus rasterisation, direct
        lighting and tonemapping: uniforms, [voxelise], allocate, light, clear HDR, apply."""
        self.m_passedTime += timeSinceLastFrame
        self.UpdatePerFrameUBO(camera)
        if not detachViewFromCameraUpdate:
            self.UpdateVolumeUBO(camera)
        self.PrepareLights()
        if self.m_indirectShadow:
            self.VoxelizeScene(timeSinceLastFrame)
        if not detachViewFromCameraUpdate:
            self.AllocateCaches()
            self.LightCachesRSM()
            if self.m_indirectSpecular:  # renderer.cpp:557-558
                self.PrepareSpecularEnvmaps()
        if hdr is None and self._hdr is not None:  # glClear(GL_COLOR_BUFFER_BIT), renderer.cpp:562
            if self._tstream is not None:
                import torch
                with torch.cuda.stream(self._tstream):
                    self._hdr.zero_()
            else:
                self._hdr.zero_()
        return self.ApplyCaches(hdr)
