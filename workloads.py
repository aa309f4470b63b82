"""The synthetic workloads of BASELINE.json (`configs`), shared by tests/ and bench.py.

Every workload is procedural (scenes/) and deterministic. Host-side arrays are
numpy in the reference encodings; uniform blocks come from the packers of
``include/drv_math.h`` through the C-ABI (``drv_pack_*``).

    C1  Cornell box 512x512, one 64x64 RSM (4k VPLs), 1 cascade x 32^3, SH1, no shadow
    C2  atrium 1920x1080, RSM 1024^2 read at LOD 3 = 128^2 (16k VPLs), 2 x 64^3 (8/16 m), SH1, unshadowed
    C3  C2 + 128^3 voxels + cone-traced shadows (LOD 2), SH2
    C4  atrium x2 at 3840x2160, 4 lights x 128^2 (64k VPLs), 4 x 128^3, SH2 + shadows
    C5  gather sweep: seeded synthetic entries x VPLs (no scene)
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import dynamicradiancevolume_b200 as drv
from dynamicradiancevolume_b200 import abi


@dataclass
class Workload:
    name: str
    width: int
    height: int
    scene: str
    scene_scale: float
    camera: drv.Camera
    lights: List[drv.Light]
    cav_cascades: int
    cav_resolution: int
    cascade_sizes: List[float]
    transition: float
    sh_order: int
    indirect_shadow: bool
    voxel_resolution: int
    max_caches: int
    # filled by build()
    constant: Optional[abi.Constant] = None
    per_frame: Optional[abi.PerFrame] = None
    volume: Optional[abi.VolumeInfo] = None
    spot_lights: List[abi.SpotLight] = field(default_factory=list)
    depth: Optional[np.ndarray] = None
    normal: Optional[np.ndarray] = None
    diffuse: Optional[np.ndarray] = None
    rsms: list = field(default_factory=list)       # [(flux, normal, depthLinSq)] level 0
    triangles: Optional[np.ndarray] = None
    bbox: tuple = ()

    @property
    def transitions(self) -> bool:
        return self.transition > 0.0

    @property
    def num_vpls(self) -> int:
        return sum(int(s.RSMReadResolution) ** 2 for s in self.spot_lights)

    def build(self, threads: int = 0, render: bool = True):
        """Pack the uniform blocks and (optionally) ray-cast the G-buffer and the RSMs."""
        from scenes import binding as scn
        geo = scn.SceneGeometry(self.scene, self.scene_scale)
        self.bbox = geo.bounding_box()
        self.triangles = geo.triangles()
        self.camera.aspect_ratio = self.width / self.height
        self.constant = drv.pack_constant(self.width, self.height, self.voxel_resolution, self.cav_resolution,
                                          self.cav_cascades, self.max_caches)
        self.per_frame = drv.pack_per_frame(self.camera, 0.0)
        self.volume = drv.pack_volume_info(self.camera, self.bbox[0], self.bbox[1], self.voxel_resolution,
                                           self.cav_resolution, self.cascade_sizes, self.transition)
        self.spot_lights = [drv.pack_spot_light(l) for l in self.lights]
        if render:
            self.depth, self.normal, self.diffuse = geo.render_gbuffer(self.per_frame, self.width, self.height, threads)
            self.rsms = [geo.render_rsm(s, threads) for s in self.spot_lights]
        return self

    def context_kwargs(self):
        return dict(max_cache_count=self.max_caches, cav_cascades=self.cav_cascades, cav_resolution=self.cav_resolution,
                    voxel_resolution=self.voxel_resolution, sh_order=self.sh_order, indirect_shadow=self.indirect_shadow,
                    cascade_transitions=self.transitions, width=self.width, height=self.height,
                    max_lights=max(1, len(self.lights)),
                    max_rsm_resolution=max([int(s.RSMRenderResolution) for s in self.spot_lights] + [16]))


def _spot(position, direction, rsm_res, read_lod, shadow_lod=2, intensity=100.0, half_angle_deg=30.0):
    return drv.Light(intensity=(intensity,) * 3, position=position, direction=direction,
                     halfAngle=half_angle_deg * math.pi / 180.0, rsmResolution=rsm_res, rsmReadLod=read_lod,
                     indirectShadowComputationLod=shadow_lod)


def cornell(width=512, height=512, rsm_res=64, read_lod=0, sh_order=1, indirect_shadow=False, cav_resolution=32,
            voxel_resolution=64, transition=0.0, max_caches=16384, shadow_lod=2) -> Workload:
    """C1 (application.cpp:51-52, 86-91 camera/light defaults). One 16 m cascade so the box fits around the camera."""
    cam = drv.Camera(position=(0.0, 2.5, 5.0), direction=(0.0, -2.5, -5.0))
    light = _spot((0.0, 1.7, 3.3), (0.0, 0.0, -1.0), rsm_res, read_lod, shadow_lod)
    return Workload("C1-cornell", width, height, "cornell", 1.0, cam, [light], 1, cav_resolution, [16.0], transition,
                    sh_order, indirect_shadow, voxel_resolution, max_caches)


def atrium(width=1920, height=1080, rsm_res=1024, read_lod=3, sh_order=1, indirect_shadow=False, cascades=2,
           cav_resolution=64, first_cascade=8.0, voxel_resolution=128, transition=2.0, max_caches=65536,
           shadow_lod=2, num_lights=1, scale=1.0, name="C2-atrium") -> Workload:
    """C2 / C3 (and C4 with scale=2, 4 lights, 4 cascades of 128^3)."""
    s = scale
    cam = drv.Camera(position=(0.0, 2.5 * s, 2.0 * s), direction=(0.0, -0.12, -1.0))
    spots = [((-1.0 * s, 6.2 * s, 3.0 * s), (0.25, -1.0, -0.55)),
             ((1.5 * s, 6.0 * s, -2.0 * s), (-0.3, -1.0, 0.2)),
             ((-3.9 * s, 3.2 * s, 0.0 * s), (1.0, -0.35, -0.3)),
             ((3.9 * s, 3.0 * s, 4.0 * s), (-1.0, -0.4, -0.5))]
    lights = [_spot(p, d, rsm_res, read_lod, shadow_lod) for p, d in spots[:num_lights]]
    sizes = [first_cascade * s * (2.0 ** i) for i in range(cascades)]
    return Workload(name, width, height, "atrium", s, cam, lights, cascades, cav_resolution, sizes, transition,
                    sh_order, indirect_shadow, voxel_resolution, max_caches)


def config(index: int, **kw) -> Workload:
    """BASELINE.json `configs[index]` (0-based)."""
    if index == 0:
        return cornell(**kw)
    if index == 1:
        return atrium(**kw)
    if index == 2:
        return atrium(sh_order=2, indirect_shadow=True, name="C3-atrium-shadow-sh2", **kw)
    if index == 3:
        args = dict(width=3840, height=2160, rsm_res=1024, read_lod=3, sh_order=2, indirect_shadow=True, cascades=4,
                    cav_resolution=128, first_cascade=4.0, num_lights=4, scale=2.0, max_caches=1 << 20,
                    name="C4-atrium-4k")
        args.update(kw)
        return atrium(**args)
    if index == 5:
        return reference_defaults(**kw)
    if index == 6:
        return dense(**kw)
    raise ValueError("config index 0..3, 5 or 6 (the sweep, index 4, has no scene: see sweep())")


def reference_defaults(**kw) -> Workload:
    """Not a BASELINE config: the settings the reference's author ran by default on the atrium —
    1920x1080 (outputwindow.cpp:59-60), one light, RSM 1024^2 read at LOD 4 = 64^2 = 4096 VPLs (scene/light.hpp:12-14),
    3 cascades x 32^3 of 4 / 8 / 16 m with 2-voxel transitions (renderer.cpp:41-48, 1184-1187), 128^3 voxels,
    at most 16384 caches (renderer.cpp:86-90), SH1, indirect shadows at LOD 2."""
    args = dict(width=1920, height=1080, rsm_res=1024, read_lod=4, sh_order=1, indirect_shadow=True, cascades=3,
                cav_resolution=32, first_cascade=4.0, voxel_resolution=128, transition=2.0, max_caches=16384,
                shadow_lod=2, name="reference-defaults")
    args.update(kw)
    return atrium(**args)


def dense(**kw) -> Workload:
    """Not a BASELINE config: configs[1] with a four times finer address volume (2 x 256^3 instead of 2 x 64^3, 3 cm
    cells in the first cascade), which allocates ~77 000 caches on the same view instead of 6 210 — the upper half of
    SURVEY 8a's "10^4-10^5 caches for real frames". bench.py carries it as `dense_workload`."""
    args = dict(cav_resolution=256, max_caches=1 << 17, name="C2-atrium-dense")
    args.update(kw)
    return atrium(**args)


def sweep(n_cache: int, n_vpl: int, seed: int = 0xD27A0001):
    """C5: (positions[n,4] f32, vpls[n_vpl] VPL_DTYPE) with the SURVEY 8d generators."""
    from scenes import binding as scn
    half = 30.0 * math.pi / 180.0
    r = int(round(math.sqrt(n_vpl)))
    val_area_factor = (2.0 * math.sin(half)) ** 2 / float(max(r, 1) ** 2)
    return scn.sweep_entries(seed, n_cache), scn.sweep_vpls(seed + 0x10000000, n_vpl, val_area_factor)


def rsm_read_level(light: abi.SpotLight) -> int:
    return int(round(math.log2(light.RSMRenderResolution / light.RSMReadResolution)))


class DeviceFrame:
    """A Workload resident on one GPU, driven through the C-ABI (``drv.Context``).

    PyTorch only moves the host arrays into device memory; every stage call goes
    to libdrv_gi. ``stream`` (a ``torch.cuda.Stream``) becomes the context's
    stream so CUDA events recorded on it bracket the kernels.
    """

    def __init__(self, wl: Workload, device: int = 0, stream=None, gather_variant: int = 0, **overrides):
        import torch
        self.torch = torch
        self.wl = wl
        self.device = device
        self.stream = stream
        kw = wl.context_kwargs()
        kw.update(overrides)
        kw.update(device=device, gather_variant=gather_variant,
                  stream=None if stream is None else stream.cuda_stream)
        self.ctx = drv.Context(**kw)
        dev = "cuda:%d" % device
        self.depth = torch.from_numpy(wl.depth).to(dev)
        self.normal = torch.from_numpy(wl.normal).to(dev)
        self.diffuse = torch.from_numpy(wl.diffuse).to(dev)
        self.rsms = [tuple(torch.from_numpy(a).to(dev) for a in r) for r in wl.rsms]
        self.tris = torch.from_numpy(np.ascontiguousarray(wl.triangles, np.float32).reshape(-1)).to(dev)
        self.out32 = torch.zeros(wl.height, wl.width, 4, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(device)
        c = self.ctx
        c.set_constant(wl.constant)
        c.set_per_frame(wl.per_frame)
        c.set_volume_info(wl.volume)
        c.set_light_count(len(wl.spot_lights))
        for i, s in enumerate(wl.spot_lights):
            c.set_spot_light(i, s)
        c.bind_gbuffer(self.depth, self.normal, self.diffuse)
        for i, r in enumerate(self.rsms):
            c.bind_rsm(i, *r)

    def prepare_inputs(self):
        """RSM mip chains (ShadowMap::PrepareRSM) and, with indirect shadows, the voxel volume + mips."""
        for i in range(len(self.rsms)):
            self.ctx.prepare_rsm(i)
        if self.wl.indirect_shadow:
            self.ctx.voxelize(self.tris, None, 1.0)

    def frame(self, out=None, fmt=abi.DRV_HDR_RGBA32F_WRITE):
        """allocate -> light -> apply (the DYN_RADIANCE_VOLUME case of Renderer::Draw)."""
        self.ctx.draw(self.out32 if out is None else out, fmt)

    def close(self):
        self.ctx.close()
