"""GPU parity, stage by stage: libdrv_gi (through the C-ABI) against the CPU oracle on the same inputs.

Gates (BASELINE.json north_star): allocation / compaction / voxel sets bit-exact; SH coefficients and
radiance within 1e-3 relative / 1e-5 absolute (``oracle.frame.close``).
"""
import ctypes as C

import numpy as np
import pytest

import dynamicradiancevolume_b200 as drv
import workloads
from dynamicradiancevolume_b200 import abi
from oracle import binding as orc
from oracle.frame import OracleFrame, close

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _frames(wl, **kw):
    wl.build()
    g = workloads.DeviceFrame(wl, **kw)
    o = OracleFrame(wl)
    return g, o


# ---------------------------------------------------------------------------- stage 1: allocation
@pytest.mark.parametrize("name,make", [
    ("cornell", lambda: workloads.cornell()),
    ("cornell-sh2", lambda: workloads.cornell(sh_order=2)),
    ("cornell-ragged", lambda: workloads.cornell(width=333, height=201)),
    ("atrium-small-transitions", lambda: workloads.atrium(width=480, height=270, rsm_res=64, read_lod=0)),
    ("atrium-3casc", lambda: workloads.atrium(width=640, height=360, rsm_res=64, read_lod=0, cascades=3,
                                              cav_resolution=32, first_cascade=4.0)),
    ("atrium-1080p", lambda: workloads.atrium(rsm_res=64, read_lod=0)),
    # a frame whose second-cascade cells are partly registered only by the threads with local x, y >= 14 of a tile
    # (the `lookUpThread == LOCAL_SIZE - 1` clause of cacheGather.comp:147-150)
    ("atrium-r16-transitions", lambda: workloads.atrium(width=160, height=90, rsm_res=32, read_lod=0, cav_resolution=16,
                                                        first_cascade=8.0, max_caches=8192)),
])
def test_allocation_bit_exact(cuda_device, name, make):
    wl = make()
    g, o = _frames(wl)
    g.ctx.allocate_caches()
    n, overflow, oob = g.ctx.active_cache_count()
    o.allocate()
    assert (n, overflow, oob) == (o.count, o.alloc["overflow"], o.alloc["oob"])
    assert n > 0
    atlas = g.ctx.read_atlas()
    # the parity object: the set of allocated linear cell ids, bit-exact after sorting
    R, Cn = wl.cav_resolution, wl.cav_cascades
    zz, yy, xx = np.nonzero(atlas)
    ids_gpu = np.sort((xx % R) + yy * R + zz * R * R + (xx // R) * R ** 3).astype(np.int32)
    ids_orc = orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.depth)
    assert np.array_equal(ids_gpu, ids_orc)
    # both assign indices in ascending cell order, so atlas and entries are identical too
    assert np.array_equal(atlas, o.alloc["atlas"])
    e = g.ctx.read_entries(n)
    assert np.array_equal(e.view(np.uint32), o.entries[:n].view(np.uint32))
    # cachePrepareLighting.comp: indirect args
    b = g.ctx.buffers()
    counter = g.ctx._read(b.counter, 16).view(np.uint32)
    assert list(counter) == [(n + 63) // 64, 1, 1, n]
    # idempotent: a second allocation of the same frame yields the same bytes
    g.ctx.allocate_caches()
    assert np.array_equal(g.ctx.read_atlas(), atlas)
    g.close()


def test_allocation_empty_gbuffer(cuda_device):
    wl = workloads.cornell(width=128, height=128).build()
    wl.depth[:] = 0.0
    g = workloads.DeviceFrame(wl)
    g.ctx.allocate_caches()
    assert g.ctx.active_cache_count() == (0, 0, 0)
    g.ctx.light_caches()  # zero entries: must be a no-op, not a crash
    g.frame()
    assert float(g.out32.abs().max()) == 0.0
    g.close()


def test_allocation_capacity_overflow(cuda_device):
    wl = workloads.cornell(max_caches=100).build()
    g = workloads.DeviceFrame(wl)
    o = OracleFrame(wl).allocate()
    g.ctx.allocate_caches()
    n, overflow, _ = g.ctx.active_cache_count()
    assert n == 100 and overflow == o.alloc["overflow"] and overflow > 0
    st = g.ctx.lib.drv_active_cache_count(g.ctx.handle, None, None, None)
    assert st == abi.DRV_ERR_CAPACITY
    assert np.array_equal(g.ctx.read_atlas(), o.alloc["atlas"])  # dropped cells stay 0
    assert np.array_equal(g.ctx.read_entries(100), o.entries[:100])
    g.frame()  # lighting + apply still run on the clamped list
    g.close()


# ---------------------------------------------------------------------------- stage 2: RSM mips + VPLs
def test_rsm_mip_chain(cuda_device):
    wl = workloads.atrium(width=64, height=64, rsm_res=256, read_lod=2).build()
    g = workloads.DeviceFrame(wl)
    g.ctx.prepare_rsm(0)
    levels = orc.rsm_mip_chain(*wl.rsms[0])
    assert len(levels) == 8  # 256 .. 2: the 1x1 level is never rendered
    for l in range(1, len(levels)):
        f, n, d = g.ctx.read_rsm_mip(0, 256, l)
        fo, no, do = levels[l]
        assert np.array_equal(f[..., :3], fo[..., :3]), "flux level %d" % l
        assert np.array_equal(d, do), "depthLinSq level %d" % l
        # normals go through cos/sin/atan2 (continuous maths): allow 1 LSB of the int16 code, modulo the phi wrap
        dn = np.abs(n.astype(np.int32) - no.astype(np.int32))
        dn[..., 0] = np.minimum(dn[..., 0], 65536 - dn[..., 0])
        assert dn.max() <= 1, "normal level %d" % l
    g.close()


@pytest.mark.parametrize("shadow", [False, True])
def test_vpl_generation(cuda_device, shadow):
    # RSM bound directly at the read resolution: VPLs do not depend on the mip normals
    wl = workloads.atrium(width=64, height=64, rsm_res=128, read_lod=0, indirect_shadow=shadow, sh_order=1)
    g, o = _frames(wl)
    g.prepare_inputs()
    o.prepare_inputs()
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    v = g.ctx.read_vpls(0, 128 * 128)
    vo = o.vpls[0]
    assert np.array_equal(v["Position"], vo["Position"])  # decision maths: bit-exact
    assert np.array_equal(v["DiscArea"], vo["DiscArea"])
    assert np.array_equal(v["Flux"], vo["Flux"])
    assert np.abs(v["Normal"] - vo["Normal"]).max() < 2e-6
    if shadow:
        b = g.ctx.read_shadow_blocks(0, 128 * 128 // 16)
        assert np.array_equal(b["AverageValPos"], o.blocks[0]["AverageValPos"])
        assert np.array_equal(b["DistToSphereRad"], o.blocks[0]["DistToSphereRad"])
    g.close()


# ---------------------------------------------------------------------------- stage 3: voxels
@pytest.mark.parametrize("scene,res", [("cornell", 64), ("atrium", 128), ("atrium", 32), ("atrium", 256),
                                       ("atrium", 512)])  # 512^3: the top of the reference's range (tweakbarsetup.cpp:185)
def test_voxelize_blend_mips_bit_exact(cuda_device, scene, res):
    wl = (workloads.cornell(indirect_shadow=True, voxel_resolution=res) if scene == "cornell" else
          workloads.atrium(width=64, height=64, rsm_res=64, read_lod=0, indirect_shadow=True, voxel_resolution=res))
    g, o = _frames(wl)
    o.prepare_inputs()
    g.ctx.voxelize(g.tris, None, 1.0)
    assert np.array_equal(g.ctx.read_voxel_target(), o.target)
    assert o.target.any()
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain)
    g.close()


def test_voxel_temporal_blend(cuda_device):
    wl = workloads.cornell(indirect_shadow=True, voxel_resolution=32)
    g, o = _frames(wl)
    res = 32
    target = orc.voxelize(wl.volume, res, wl.triangles)
    vol = np.zeros(res ** 3, np.uint8)
    for k in (7, 100, 255, 30):  # adaption = k/255 per frame (voxelblend.comp:16)
        g.ctx.voxelize(g.tris, None, k / 255.0)
        orc.voxel_blend(vol, target, res, k / 255.0)
        assert np.array_equal(g.ctx.read_voxel_chain(), orc.voxel_chain(vol, res))
    # adaption 0 is a no-op (voxelization.cpp:100)
    before = g.ctx.read_voxel_chain()
    g.ctx.voxelize(None, None, 0.0)
    assert np.array_equal(g.ctx.read_voxel_chain(), before)
    g.close()


def test_voxelize_transformed_entities(cuda_device):
    """Two entities with different world matrices, added with CLEAR / FINISH flags."""
    torch = _torch()
    wl = workloads.cornell(indirect_shadow=True, voxel_resolution=64).build()
    g = workloads.DeviceFrame(wl)
    tris = wl.triangles
    half = len(tris) // 2
    w1 = np.eye(4, dtype=np.float32)
    w2 = np.eye(4, dtype=np.float32)
    w2[:3, :3] *= 0.5
    w2[:3, 3] = (0.3, 1.0, -0.2)
    t1 = torch.from_numpy(tris[:half].reshape(-1).copy()).cuda()
    t2 = torch.from_numpy(tris[half:].reshape(-1).copy()).cuda()
    torch.cuda.synchronize()
    g.ctx.voxelize(t1, w1.ravel().tolist(), 1.0, abi.DRV_VOXELIZE_CLEAR)
    g.ctx.voxelize(t2, w2.ravel().tolist(), 1.0, abi.DRV_VOXELIZE_FINISH)
    target = orc.voxelize(wl.volume, 64, tris[:half], w1)
    target = orc.voxelize(wl.volume, 64, tris[half:], w2, target)
    assert np.array_equal(g.ctx.read_voxel_target(), target)
    g.close()


# ---------------------------------------------------------------------------- stage 4: gather
def _sweep_ctx(n_cache, n_vpl, sh_order, variant, rsm_cap=512):
    torch = _torch()
    pos, vpls = workloads.sweep(n_cache, n_vpl)
    ctx = drv.Context(max_cache_count=max(n_cache, 64), cav_cascades=1, cav_resolution=8, voxel_resolution=16,
                      sh_order=sh_order, indirect_shadow=False, cascade_transitions=False, width=16, height=16,
                      max_lights=2, max_rsm_resolution=rsm_cap, gather_variant=variant)
    cb = drv.pack_constant(16, 16, 16, 8, 1, max(n_cache, 64))
    ctx.set_constant(cb)
    ctx.set_light_count(1)
    p = torch.from_numpy(pos).cuda()
    torch.cuda.synchronize()
    ctx.set_synthetic_entries(p)
    ctx.set_vpls(0, vpls.ctypes.data, n_vpl)
    return ctx, cb, pos, vpls


def _sweep_oracle(cb, pos, vpls_list, sh_order, fp64=False):
    stride = abi.entry_stride(sh_order) // 4
    e = np.zeros((len(pos), stride), np.float32)
    e[:, :3] = pos[:, :3]
    lights = []
    for v in vpls_list:
        s = abi.SpotLight()
        s.RSMReadResolution = int(round(len(v) ** 0.5))
        assert s.RSMReadResolution ** 2 == len(v)
        s.IndirectShadowComputationSampleInterval = 1
        lights.append(s)
    vi = abi.VolumeInfo()
    orc.light_caches(cb, vi, lights, vpls_list, None, None, e, 0, len(pos), sh_order, False, fp64)
    return e


@pytest.mark.parametrize("variant", [0, 1, 3, 4, 7, 8, 9, 11, 12, 20, 21, 22, 23, 24, 25, 26, 27, 28, 30, 31, 32, 33, 34,
                                     35, 36, 37, 38, 39, 40])
@pytest.mark.parametrize("sh_order", [1, 2])
@pytest.mark.parametrize("n_cache,n_vpl", [(1000, 4096), (70000, 1024), (37, 2500), (5000, 16384), (129, 9), (300000, 100)])
def test_gather_unshadowed_variants(cuda_device, variant, sh_order, n_cache, n_vpl):
    ctx, cb, pos, vpls = _sweep_ctx(n_cache, n_vpl, sh_order, variant)
    ctx.light_caches()
    e = ctx.read_entries(n_cache)
    eo = _sweep_oracle(cb, pos, [vpls], sh_order)
    assert np.array_equal(e[:, :4], eo[:, :4])
    ok, ratio = close(e[:, 4:], eo[:, 4:])
    assert ok, "worst |err|/tol = %.3f" % ratio
    assert np.abs(eo[:, 4:]).max() > 0
    # deterministic: no float atomics anywhere, so a second run adds exactly the same amounts
    ctx.set_synthetic_entries(ctx._keep_syn)
    ctx.light_caches()
    assert np.array_equal(ctx.read_entries(n_cache), e)
    ctx.close()


@pytest.mark.parametrize("pattern", ["runs", "sparse", "all-dead", "tail"])
def test_gather_drops_zero_flux_vpls(cuda_device, pattern):
    """Live-VPL compaction (rsm.cu): VPLs with zero flux are removed before the gather, the order of the rest is
    kept, and the result equals the oracle's walk over the full list (ragged runs, chunk boundaries, empty list)."""
    n_cache, n_vpl = 777, 4096
    pos, vpls = workloads.sweep(n_cache, n_vpl)
    vpls = vpls.copy()
    k = np.arange(n_vpl)
    dead = {"runs": (k // 300) % 2 == 1, "sparse": (k * 2654435761 % 7) != 0, "all-dead": np.ones(n_vpl, bool),
            "tail": k >= 257}[pattern]
    vpls["Flux"][dead] = 0.0
    vpls["Flux"][dead & (k % 2 == 0), 1] = -0.0  # negative zero is zero too
    for sh_order in (1, 2):
        ctx, cb, _, _ = _sweep_ctx(n_cache, n_vpl, sh_order, 0)
        ctx.set_vpls(0, vpls.ctypes.data, n_vpl)
        ctx.light_caches()
        assert ctx.live_vpl_counts()[0] == int((~dead).sum())
        e = ctx.read_entries(n_cache)
        eo = _sweep_oracle(cb, pos, [vpls], sh_order)
        ok, ratio = close(e[:, 4:], eo[:, 4:])
        assert ok, "worst |err|/tol = %.3f" % ratio
        assert (np.abs(eo[:, 4:]).max() > 0) == (pattern != "all-dead")
        ctx.close()


@pytest.mark.parametrize("variant", [0, 20, 26, 28, 30, 35, 37])
def test_gather_accumulates_over_lights_and_calls(cuda_device, variant):
    """`entry.SH += acc` per light (cacheLightingRSM.comp:358-373): two lights, then a second call."""
    ctx, cb, pos, vpls = _sweep_ctx(3000, 4096, 2, variant)
    _, vpls2 = workloads.sweep(1, 1024, seed=77)
    ctx.set_light_count(2)
    ctx.set_vpls(1, vpls2.ctypes.data, 1024)
    ctx.light_caches()
    e1 = ctx.read_entries(3000)
    eo = _sweep_oracle(cb, pos, [vpls, vpls2], 2)
    ok, ratio = close(e1[:, 4:], eo[:, 4:])
    assert ok, ratio
    ctx.light_caches()  # no re-allocation in between: results add up
    e2 = ctx.read_entries(3000)
    ok, ratio = close(e2[:, 4:], 2.0 * eo[:, 4:].astype(np.float64))
    assert ok, ratio
    ctx.close()


def test_gather_sh1_is_prefix_of_sh2_and_linear_in_flux(cuda_device):
    """Size-independent properties at a BASELINE-sized problem (64k entries x 16k VPLs)."""
    n_cache, n_vpl = 65536, 16384
    ctx1, cb, pos, vpls = _sweep_ctx(n_cache, n_vpl, 1, 0)
    ctx1.light_caches()
    e1 = ctx1.read_entries(n_cache)
    ctx1.close()
    ctx2, _, _, _ = _sweep_ctx(n_cache, n_vpl, 2, 0)
    ctx2.light_caches()
    e2 = ctx2.read_entries(n_cache)
    ok, ratio = close(e1[:, 4:16], e2[:, 4:16])
    assert ok, ratio
    # linearity: scaling every VPL's flux by 4 (exact in binary) scales every coefficient by exactly 4
    v4 = vpls.copy()
    v4["Flux"] *= 4.0
    ctx2.set_synthetic_entries(ctx2._keep_syn)
    ctx2.set_vpls(0, v4.ctypes.data, n_vpl)
    ctx2.light_caches()
    e4 = ctx2.read_entries(n_cache)
    assert np.array_equal(e4[:, 4:], 4.0 * e2[:, 4:])
    # spot-check a subsample against the oracle (every 257th entry)
    sub = np.arange(0, n_cache, 257)
    eo = _sweep_oracle(cb, pos[sub], [vpls], 2)
    ok, ratio = close(e2[sub, 4:], eo[:, 4:])
    assert ok, ratio
    ctx2.close()


@pytest.mark.parametrize("sh_order", [1, 2])
@pytest.mark.parametrize("variant", [0])
def test_gather_cone_traced_shadows(cuda_device, sh_order, variant):
    wl = workloads.cornell(sh_order=sh_order, indirect_shadow=True, voxel_resolution=64)
    g, o = _frames(wl, gather_variant=variant)
    g.prepare_inputs()
    o.prepare_inputs()
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    o.allocate()
    o.light()
    n = o.count
    e = g.ctx.read_entries(n)
    ok, ratio = close(e[:, 4:], o.entries[:n, 4:])
    assert ok, "worst |err|/tol = %.3f" % ratio
    # shadows must actually bite: compare with the unshadowed result
    wl2 = workloads.cornell(sh_order=sh_order, indirect_shadow=False).build()
    o2 = OracleFrame(wl2).prepare_inputs().allocate()
    o2.light()
    assert np.abs(o2.entries[:n, 7] - o.entries[:n, 7]).max() > 1e-3
    g.close()


@pytest.mark.parametrize("shrink", [0.35, 0.6])
def test_cones_that_leave_the_voxel_volume(cuda_device, shrink):
    """The voxel volume covers only the middle of the scene, so caches and VAL blocks lie outside it and their cones
    start, end or run outside: the march's clamp-to-edge path (and the warps that mix it with the clamp-free one)
    against the oracle's sampler (clamp to edge)."""
    import dynamicradiancevolume_b200 as drv
    wl = workloads.cornell(sh_order=1, indirect_shadow=True, voxel_resolution=32).build()
    lo, hi = np.asarray(wl.bbox[0], np.float64), np.asarray(wl.bbox[1], np.float64)
    mid, half = (lo + hi) / 2, (hi - lo) / 2 * shrink
    wl.volume = drv.pack_volume_info(wl.camera, tuple(mid - half), tuple(mid + half), wl.voxel_resolution,
                                     wl.cav_resolution, wl.cascade_sizes, wl.transition)
    g, o = workloads.DeviceFrame(wl), OracleFrame(wl)
    g.prepare_inputs()
    o.prepare_inputs()
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain) and o.chain.any()
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    o.allocate()
    o.light()
    n = o.count
    pos = o.entries[:n, :3]
    vmin, vmax = np.asarray(wl.volume.VolumeWorldMin[:3]), np.asarray(wl.volume.VolumeWorldMax[:3])
    outside = ((pos < vmin) | (pos > vmax)).any(axis=1)
    assert outside.any() and not outside.all()  # both kinds of cone in one frame
    e = g.ctx.read_entries(n)
    ok, ratio = close(e[:, 4:], o.entries[:n, 4:])
    assert ok, "worst |err|/tol = %.3f" % ratio
    g.close()


def test_cone_trace_extremes(cuda_device):
    """Empty volume => shadowing 1 (equals the unshadowed gather); full volume => every SH coefficient 0."""
    torch = _torch()
    wl = workloads.cornell(sh_order=1, indirect_shadow=True, voxel_resolution=32)
    g, o = _frames(wl)
    for i in range(len(g.rsms)):
        g.ctx.prepare_rsm(i)
    g.ctx.voxelize(None, None, 1.0)  # nothing rasterised: empty volume
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    n = g.ctx.active_cache_count()[0]
    e_empty = g.ctx.read_entries(n)
    wl2 = workloads.cornell(sh_order=1, indirect_shadow=False).build()
    o2 = OracleFrame(wl2).prepare_inputs().allocate()
    o2.light()
    ok, ratio = close(e_empty[:, 4:], o2.entries[:n, 4:])
    assert ok, ratio
    full = torch.full((32 ** 3,), 255, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    g.ctx.set_voxel_volume(full)  # installs level 0, rebuilds mips + gather records
    assert np.all(g.ctx.read_voxel_chain() == 255)
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    e_full = g.ctx.read_entries(n)
    assert np.abs(e_full[:, 4:]).max() == 0.0
    g.close()


@pytest.mark.parametrize("variant", [0, 7, 8, 12, 20, 21, 25, 26, 27, 28, 30, 31, 35, 37])
def test_shadowed_gather_variants_agree(cuda_device, variant):
    """Packed (default) and scalar pair kernels reading the same visibility table."""
    wl = workloads.atrium(width=320, height=180, rsm_res=64, read_lod=0, sh_order=2, indirect_shadow=True,
                          voxel_resolution=64, shadow_lod=1)
    g, o = _frames(wl, gather_variant=variant)
    g.prepare_inputs()
    o.prepare_inputs()
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    o.allocate()
    o.light()
    e = g.ctx.read_entries(o.count)
    ok, ratio = close(e[:, 4:], o.entries[:o.count, 4:])
    assert ok, "worst |err|/tol = %.3f" % ratio
    g.close()


@pytest.mark.parametrize("shadow_lod", [0, 3, 4])
def test_cone_trace_other_sample_intervals(cuda_device, shadow_lod):
    """Shadow LOD 0 (one cone per VPL), 3 (interval 64) and 4 (interval 256: a block spans two 128-VPL tiles)."""
    wl = workloads.cornell(width=128, height=128, rsm_res=32, sh_order=1, indirect_shadow=True, voxel_resolution=32,
                           shadow_lod=shadow_lod)
    g, o = _frames(wl)
    g.prepare_inputs()
    o.prepare_inputs()
    g.ctx.allocate_caches()
    g.ctx.light_caches()
    o.allocate()
    o.light()
    e = g.ctx.read_entries(o.count)
    ok, ratio = close(e[:, 4:], o.entries[:o.count, 4:])
    assert ok, "worst |err|/tol = %.3f" % ratio
    g.close()


def test_set_voxel_volume_builds_the_mip_chain(cuda_device):
    rng = np.random.default_rng(5)
    vol = (rng.random(32 ** 3) < 0.2).astype(np.uint8) * 255
    wl = workloads.cornell(indirect_shadow=True, voxel_resolution=32).build()
    g = workloads.DeviceFrame(wl)
    g.ctx.set_voxel_volume(vol)  # host pointer
    assert np.array_equal(g.ctx.read_voxel_chain(), orc.voxel_chain(vol, 32))
    g.close()


# ---------------------------------------------------------------------------- stage 5: apply
@pytest.mark.parametrize("sh_order", [1, 2])
@pytest.mark.parametrize("transition", [0.0, 2.0])
def test_apply_isolated(cuda_device, sh_order, transition):
    """Apply alone: the oracle's lit entries are installed into the context's buffer."""
    torch = _torch()
    wl = workloads.atrium(width=480, height=270, rsm_res=64, read_lod=0, sh_order=sh_order, transition=transition)
    g, o = _frames(wl)
    o.prepare_inputs()
    img_o = o.frame()
    g.ctx.allocate_caches()
    n = g.ctx.active_cache_count()[0]
    assert n == o.count
    g.ctx.entries_tensor()[:n].copy_(torch.from_numpy(o.entries[:n]).cuda())
    torch.cuda.synchronize()
    g.ctx.apply_caches(g.out32, abi.DRV_HDR_RGBA32F_WRITE)
    torch.cuda.synchronize()
    img = g.out32.cpu().numpy()
    assert np.array_equal(img[..., 3], img_o[..., 3])  # same pixels shaded / discarded
    ok, ratio = close(img[..., :3], img_o[..., :3])
    assert ok, "worst |err|/tol = %.3f" % ratio
    assert img_o[..., :3].max() > 1e-3
    g.close()


def test_apply_rows_bands_equal_full_pass(cuda_device):
    """drv_apply_caches_rows over disjoint row bands (sort-first sharding) == one full-screen pass."""
    torch = _torch()
    wl = workloads.cornell(width=200, height=173, sh_order=2)
    g, o = _frames(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    ref = g.out32.clone()
    g.out32.fill_(-1.0)
    torch.cuda.synchronize()
    for y0, y1 in ((0, 1), (1, 64), (64, 65), (65, 170), (170, 173), (173, 400)):
        g.ctx.apply_caches_rows(g.out32, abi.DRV_HDR_RGBA32F_WRITE, y0, y1)
    torch.cuda.synchronize()
    assert torch.equal(g.out32, ref)
    g.close()


def test_apply_additive_rgba16f(cuda_device):
    """Reference blend state: GL_ONE/GL_ONE into RGBA16F, alpha untouched (renderer.cpp:119,480)."""
    torch = _torch()
    wl = workloads.cornell(width=256, height=256)
    g, o = _frames(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    ref = g.out32.cpu().numpy()
    hdr = torch.full((256, 256, 4), 0.25, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    g.ctx.apply_caches(hdr, abi.DRV_HDR_RGBA16F_ADD)
    torch.cuda.synchronize()
    out = hdr.float().cpu().numpy()
    expect = (np.float32(0.25) + ref[..., :3]).astype(np.float16).astype(np.float32)
    shaded = ref[..., 3] > 0
    assert np.array_equal(out[..., :3][shaded], expect[shaded])
    assert np.all(out[..., 3] == 0.25)
    assert np.all(out[..., :3][~shaded] == 0.25)
    g.close()
