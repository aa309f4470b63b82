"""Known-answer tests that pin the CPU oracle (oracle/) — CPU only.

The reference ships no golden vectors for this path and cannot run here ("parity unpinned"), so the
oracle is pinned by (a) hand-derived closed-form cases, (b) independent numpy restatements of the cited shader
lines in float64, and (c) the committed regression fixtures in tests/golden/ (tests/test_golden.py).
"""
import math

import numpy as np
import pytest

import dynamicradiancevolume_b200 as drv
import workloads
from dynamicradiancevolume_b200 import abi
from oracle import binding as orc
from oracle.frame import OracleFrame


# ---------------------------------------------------------------------------------- encodings
def test_half_conversions_match_ieee_binary16():
    codes = np.arange(65536, dtype=np.uint16)
    ref = codes.view(np.float16).astype(np.float32)
    got = np.array([orc.half_to_float(int(c)) for c in codes[::7]], np.float32)
    want = ref[::7]
    assert np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-7, 1e-4, 1.0, 300.0, 7e4)])
    vals = np.concatenate([vals, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 5.96e-8, 2.98e-8, 2.9802325e-8,
                                           1.0009765625, 1.00048828125, 1.00146484375], np.float32)])
    with np.errstate(over="ignore"):
        want16 = vals.astype(np.float16).view(np.uint16)
    got16 = np.array([orc.float_to_half(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got16, want16)  # round-to-nearest-even, overflow to inf, subnormals


def test_srgb_and_morton_and_normals():
    for v in (0, 1, 10, 11, 128, 255):
        c = v / 255.0
        want = c / 12.92 if c <= 0.04045 else ((c + 0.055) / 1.055) ** 2.4
        assert abs(orc.srgb8_to_linear(v) - want) < 1e-7
    for k in (0, 1, 2, 3, 0b1101, 0xFFFF, 0xAAAA5555, 12345678):
        x = sum(((k >> (2 * i)) & 1) << i for i in range(16))
        y = sum(((k >> (2 * i + 1)) & 1) << i for i in range(16))
        assert orc.morton_decode(k) == (x, y)  # cacheLightingRSM.comp:46-62
    # PackNormal16I / UnpackNormal16I (utils.glsl:44-89): round trip within the 16-bit quantisation
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        px, py = orc.pack_normal16i(n)
        back = orc.unpack_normal16i(px, py)
        assert np.abs(back - n).max() < 3e-4
    assert orc.pack_normal16i((0.0, 0.0, 1.0)) == (0, 32767)       # n.z = 1 would be 32768: clamped
    assert orc.pack_normal16i((0.0, 1.0, 0.0)) == (16384, 0)       # x == 0 special case: sign(y) * pi/2
    assert orc.pack_normal16i((0.0, -1.0, 0.0)) == (-16384, 0)
    assert orc.pack_normal16i((-1.0, 0.0, 0.0))[0] in (32767, -32768)


# ---------------------------------------------------------------------------------- stage 1
def _one_pixel_frame(x, y, depth_value=0.5, res=32, transitions=False):
    wl = workloads.cornell(width=64, height=64, rsm_res=16, cav_resolution=res,
                           transition=2.0 if transitions else 0.0).build(render=False)
    depth = np.zeros((64, 64), np.float32)
    depth[y, x] = depth_value
    return wl, depth


def _world_pos(wl, x, y, d):
    """cacheGather.comp:113-119 in float32 numpy, the policy's operation order."""
    f = np.float32
    ndc = np.array([f(f(f(x + 0.5) / f(wl.width)) * f(2)) - f(1), f(f(f(y + 0.5) / f(wl.height)) * f(2)) - f(1), f(d), f(1)], f)
    m = np.array(list(wl.per_frame.InverseViewProjection), f).reshape(4, 4)
    w4 = np.zeros(4, f)
    for j in range(4):
        w4[j] = f(f(f(m[j, 0] * ndc[0]) + f(m[j, 1] * ndc[1])) + f(m[j, 2] * ndc[2])) + f(m[j, 3] * ndc[3])
    return (w4[:3] / w4[3]).astype(f)


def test_one_pixel_allocates_its_eight_corner_cells():
    # local id (3,5) of tile (1,2): all three compared neighbours hold -1 => the pixel fires
    wl, depth = _one_pixel_frame(16 + 3, 32 + 5)
    a = orc.allocate_caches(wl.constant, wl.per_frame, wl.volume, False, depth, 1, 64)
    assert a["count"] == 8 and a["overflow"] == 0 and a["oob"] == 0
    c0 = wl.volume.AddressVolumeCascades[0]
    wp = _world_pos(wl, 19, 37, 0.5)
    cell = np.clip(((wp - np.array(list(c0.Min), np.float32)) / np.float32(c0.WorldVoxelSize)).astype(np.int32), 0, 31)
    R = 32
    want = sorted(int((cell[0] + dx) + (cell[1] + dy) * R + (cell[2] + dz) * R * R)
                  for dx in (0, 1) for dy in (0, 1) for dz in (0, 1))
    ids = orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, False, depth)
    assert list(ids) == want
    # entries in ascending cell order; Position = cell * voxel + Min (cacheGather.comp:65); atlas = index + 1
    for i, cid in enumerate(want):
        x, yy, z = cid % R, (cid // R) % R, cid // (R * R)
        pos = np.array([x, yy, z], np.float32) * np.float32(c0.WorldVoxelSize) + np.array(list(c0.Min), np.float32)
        assert np.array_equal(a["entries"][i, :3], pos)
        assert a["atlas"][z, yy, x] == i + 1
        assert not a["entries"][i, 3:].any()
    assert np.count_nonzero(a["atlas"]) == 8
    assert (a["counter"].NumCacheLightingThreadGroupsX, a["counter"].NumCacheLightingThreadGroupsY,
            a["counter"].NumCacheLightingThreadGroupsZ, a["counter"].TotalLightCacheCount) == (1, 1, 1, 8)


def test_tile_edge_quirk_of_the_trigger_predicate():
    """SURVEY B.1 (cacheGather.comp:142-150): a lone pixel in column 0 / row 0 of a 16x16 tile compares against
    its own slot and never allocates — except local (0,0), which always does."""
    for (x, y, want) in [(16, 32 + 5, 0), (16 + 3, 32, 0), (16, 32, 8), (16 + 1, 32 + 1, 8), (31, 47, 8)]:
        wl, depth = _one_pixel_frame(x, y)
        n = len(orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, False, depth))
        assert n == want, (x, y, n)


def test_depth_thresholds_and_dedupe():
    wl, depth = _one_pixel_frame(19, 37, depth_value=0.0001)  # not > 1e-4: no cache (cacheGather.comp:109)
    assert len(orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, False, depth)) == 0
    # a 2x2 block of pixels in the same cell still yields 8 caches (idempotent marking replaces the CAS lock)
    wl, depth = _one_pixel_frame(19, 37)
    depth[37:39, 19:21] = 0.5
    ids = orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, False, depth)
    assert len(ids) in (8, 12, 18, 27) and len(ids) == len(set(ids.tolist()))


# ---------------------------------------------------------------------------------- stage 4
def _constant():
    return drv.pack_constant(16, 16, 16, 8, 1, 64)


def _light(n_vpl):
    s = abi.SpotLight()
    s.RSMReadResolution = int(round(math.sqrt(n_vpl)))
    s.IndirectShadowComputationSampleInterval = 1
    return s


def test_one_cache_one_vpl_closed_form_sh():
    cb = _constant()
    vpl = np.zeros(1, abi.VPL_DTYPE)
    vpl["Position"] = (0.0, 0.0, 2.0)
    vpl["Normal"] = (0.0, 0.0, -1.0)
    vpl["Flux"] = (1.0, 2.0, 3.0)
    vpl["DiscArea"] = 0.5
    for order in (1, 2):
        e = np.zeros((1, abi.entry_stride(order) // 4), np.float32)
        orc.light_caches(cb, abi.VolumeInfo(), [_light(1)], [vpl], None, None, e, 0, 1, order, False)
        rad = np.array([1.0, 2.0, 3.0]) / (4.0 + 0.5)  # toVal = (0,0,1), d^2 = 4, cos = 1
        np.testing.assert_allclose([e[0, 7], e[0, 11], e[0, 15]], cb.ShEvaFactor0 * rad, rtol=1e-6)  # SH00
        np.testing.assert_allclose(e[0, 8:11], cb.ShEvaFactor1 * rad, rtol=1e-6)                     # SH10 += f1 z rad
        assert not e[0, 4:7].any() and not e[0, 12:15].any()                                         # y = x = 0
        if order == 2:
            np.testing.assert_allclose([e[0, 19], e[0, 23], e[0, 27]], cb.ShEvaFactor20 * 2.0 * rad, rtol=1e-6)  # 3z^2-1
            assert not e[0, 16:19].any() and not e[0, 20:23].any() and not e[0, 24:27].any() and not e[0, 28:31].any()
    # a VPL facing away contributes nothing (saturate, :256)
    vpl["Normal"] = (0.0, 0.0, 1.0)
    e = np.zeros((1, 16), np.float32)
    orc.light_caches(cb, abi.VolumeInfo(), [_light(1)], [vpl], None, None, e, 0, 1, 1, False)
    assert not e[0, 4:].any()


def _numpy_gather(cb, pos, vpls, order):
    """cacheLightingRSM.comp:249-277 restated independently in float64 numpy."""
    P = pos[:, None, :3].astype(np.float64)
    t = vpls["Position"][None].astype(np.float64) - P
    d2 = (t * t).sum(-1)
    t = t / np.sqrt(d2)[..., None]
    cosv = np.clip(-(vpls["Normal"][None].astype(np.float64) * t).sum(-1), 0.0, 1.0)
    s = cosv / (d2 + vpls["DiscArea"][None].astype(np.float64))
    rad = vpls["Flux"][None].astype(np.float64) * s[..., None]
    x, y, z = t[..., 0:1], t[..., 1:2], t[..., 2:3]
    out = np.zeros((len(pos), 32 if order == 2 else 16))
    out[:, 4:7] = -(cb.ShEvaFactor1 * y * rad).sum(1)
    out[:, 8:11] = (cb.ShEvaFactor1 * z * rad).sum(1)
    out[:, 12:15] = -(cb.ShEvaFactor1 * x * rad).sum(1)
    out[:, [7, 11, 15]] = (cb.ShEvaFactor0 * rad).sum(1)
    if order == 2:
        f2 = cb.ShEvaFactor2n2_p1_n1
        out[:, 16:19] = -(f2 * x * y * rad).sum(1)
        out[:, 20:23] = (f2 * y * z * rad).sum(1)
        out[:, 24:27] = (f2 * x * z * rad).sum(1)
        out[:, 28:31] = (cb.ShEvaFactor2p2 * (x * x - y * y) * rad).sum(1)
        out[:, [19, 23, 27]] = (cb.ShEvaFactor20 * (3 * z * z - 1) * rad).sum(1)
    return out


@pytest.mark.parametrize("order", [1, 2])
def test_gather_against_independent_float64_numpy(order):
    pos, vpls = workloads.sweep(96, 1024)
    cb = _constant()
    want = _numpy_gather(cb, pos, vpls, order)
    for fp64, rtol, atol in ((True, 2e-5, 1e-9), (False, 3e-4, 2e-7)):
        e = np.zeros((96, abi.entry_stride(order) // 4), np.float32)
        e[:, :3] = pos[:, :3]
        orc.light_caches(cb, abi.VolumeInfo(), [_light(1024)], [vpls], None, None, e, 0, 96, order, False, fp64)
        np.testing.assert_allclose(e[:, 4:], want[:, 4:], rtol=rtol, atol=atol)
    # the SH1 result equals the first four coefficient groups of the SH2 result bit for bit
    e1 = np.zeros((96, 16), np.float32); e1[:, :3] = pos[:, :3]
    e2 = np.zeros((96, 32), np.float32); e2[:, :3] = pos[:, :3]
    orc.light_caches(cb, abi.VolumeInfo(), [_light(1024)], [vpls], None, None, e1, 0, 96, 1, False)
    orc.light_caches(cb, abi.VolumeInfo(), [_light(1024)], [vpls], None, None, e2, 0, 96, 2, False)
    assert np.array_equal(e1[:, 4:16], e2[:, 4:16])


def test_gather_entry_range_and_accumulation():
    """`id < TotalLightCacheCount` guard and `entry.SH += acc` (cacheLightingRSM.comp:341-374)."""
    pos, vpls = workloads.sweep(40, 256)
    cb = _constant()
    e = np.zeros((40, 16), np.float32); e[:, :3] = pos[:, :3]
    orc.light_caches(cb, abi.VolumeInfo(), [_light(256)], [vpls], None, None, e, 10, 20, 1, False)
    assert not e[:10, 4:].any() and not e[30:, 4:].any() and e[10:30, 7].all()
    once = e.copy()
    orc.light_caches(cb, abi.VolumeInfo(), [_light(256)], [vpls], None, None, e, 10, 20, 1, False)
    np.testing.assert_allclose(e[10:30, 4:], 2 * once[10:30, 4:], rtol=1e-6)
    # ragged thread count > entries
    e3 = np.zeros((3, 16), np.float32); e3[:, :3] = pos[:3, :3]
    orc.light_caches(cb, abi.VolumeInfo(), [_light(256)], [vpls], None, None, e3, 0, 3, 1, False, False, 16)
    assert np.array_equal(e3[:, 4:], once[:0, 4:]) or e3[:, 7].all()


# ---------------------------------------------------------------------------------- voxels + cones
def _volume(res=16):
    vi = abi.VolumeInfo()
    vi.VolumeWorldMin[:] = (0.0, 0.0, 0.0)
    vi.VolumeWorldMax[:] = (float(res),) * 3
    vi.VoxelSizeInWorld = 1.0
    return vi


def test_voxel_sampler_known_answers():
    res = 8
    lvl0 = np.zeros((res, res, res), np.uint8)
    lvl0[2, 3, 4] = 255  # z, y, x
    chain = orc.voxel_chain(lvl0.reshape(-1), res)
    assert len(chain) == 512 + 64 + 8 + 1
    centre = ((4 + 0.5) / res, (3 + 0.5) / res, (2 + 0.5) / res)
    assert orc.sample_voxel(chain, res, centre, 0.0) == 1.0                  # texel centre: exact texel
    assert orc.sample_voxel(chain, res, centre, -3.0) == 1.0                 # lod < 0 clamps to 0
    half = ((4 + 1.0) / res, (3 + 0.5) / res, (2 + 0.5) / res)
    assert abs(orc.sample_voxel(chain, res, half, 0.0) - 0.5) < 1e-6         # halfway to an empty neighbour
    # mips: mean of 8 children, rounded to UNORM8 (voxelmipmap.comp:11-12): 255/8 = 31.875 -> 32
    l1 = chain[512:576].reshape(4, 4, 4)
    assert l1[1, 1, 2] == 32 and np.count_nonzero(l1) == 1
    assert chain[576:584].reshape(2, 2, 2)[0, 0, 1] == 4 and chain[584] == 1  # 32/8 = 4, 4/8 = 0.5 -> 1 (ties up)
    c1 = ((2 + 0.5) / 4, (1 + 0.5) / 4, (1 + 0.5) / 4)
    a0, a1 = orc.sample_voxel(chain, res, c1, 0.0), orc.sample_voxel(chain, res, c1, 1.0)
    assert abs(a1 - 32 / 255.0) < 1e-7
    assert abs(orc.sample_voxel(chain, res, c1, 0.25) - (0.75 * a0 + 0.25 * a1)) < 1e-6  # mip-linear
    assert abs(orc.sample_voxel(chain, res, c1, 99.0) - 1 / 255.0) < 1e-7                # clamps to the 1^3 level
    assert orc.sample_voxel(chain, res, (-5.0, 0.5, 0.5), 0.0) == 0.0                    # clamp to edge


def test_cone_trace_empty_full_and_wall():
    res = 16
    vi = _volume(res)
    blk = np.zeros(1, abi.SHADOW_BLOCK_DTYPE)
    blk["AverageValPos"] = (14.5, 8.5, 8.5)
    blk["DistToSphereRad"] = 0.05
    pos = (1.5, 8.5, 8.5)
    empty = orc.voxel_chain(np.zeros(res ** 3, np.uint8), res)
    assert orc.cone_trace(vi, empty, res, pos, blk) == 1.0           # empty volume => unshadowed
    full = np.full(len(empty), 255, np.uint8)
    assert orc.cone_trace(vi, full, res, pos, blk) == 0.0            # full volume => fully shadowed
    wall = np.zeros((res, res, res), np.uint8)
    wall[:, :, 8] = 255                                              # a solid x = 8 slab between cache and light
    assert orc.cone_trace(vi, orc.voxel_chain(wall.reshape(-1), res), res, pos, blk) < 0.05
    wall[:] = 0
    wall[:, :, 15] = 255                                             # geometry behind the light: not reached
    assert orc.cone_trace(vi, orc.voxel_chain(wall.reshape(-1), res), res, pos, blk) > 0.9


def test_voxel_blend_integer_form_equals_the_float_shader():
    """voxelblend.comp:16 evaluated in float then stored to UNORM8 == the oracle's integer form, for every
    (old, target) pair and a spread of adaption steps."""
    old = np.repeat(np.arange(256, dtype=np.uint8), 256)
    tgt = np.tile(np.arange(256, dtype=np.uint8), 256)
    pad = 4  # 64^3 = 4 x 65536
    for k in (1, 2, 3, 7, 50, 127, 128, 254, 255):
        vol = np.tile(old, pad).copy()
        orc.voxel_blend(vol, np.tile(tgt, pad), 64, k / 255.0)
        o = old.astype(np.float64) / 255.0
        t = tgt.astype(np.float64) / 255.0
        v = o + np.sign(t - o) * (k / 255.0)
        want = np.floor(np.clip(v, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
        assert np.array_equal(vol[:65536], want), k


def test_voxelizer_axis_aligned_quad_fills_one_slab():
    """A floor-sized quad at y = 3.5 voxels fills exactly the y = 3 slab (plus nothing else): hand-checkable."""
    res = 16
    vi = _volume(res)
    y = 3.5
    quad = np.array([[2, y, 2, 14, y, 2, 14, y, 14], [2, y, 2, 14, y, 14, 2, y, 14]], np.float32)
    vol = orc.voxelize(vi, res, quad).reshape(res, res, res)  # z, y, x
    assert vol[:, 3, :].any() and not vol[:, :3, :].any() and not vol[:, 4:, :].any()
    xs = np.nonzero(vol[:, 3, :].any(0))[0]
    zs = np.nonzero(vol[:, 3, :].any(1))[0]
    # conservative: the covered range includes every voxel the quad touches and at most one ring more
    assert xs.min() in (1, 2) and xs.max() in (13, 14) and zs.min() in (1, 2) and zs.max() in (13, 14)
    assert vol[2:14, 3, 2:14].all()


def test_rsm_downsample_against_numpy_half_arithmetic():
    rng = np.random.default_rng(3)
    r = 8
    flux = np.zeros((r, r, 4), np.float16)
    flux[..., :3] = rng.random((r, r, 3)) * 0.01
    d = (rng.random((r, r)) * 10 + 1).astype(np.float16)
    depth = np.stack([d, (d.astype(np.float32) ** 2).astype(np.float16)], -1)
    normal = np.zeros((r, r, 2), np.int16)
    for yy in range(r):
        for xx in range(r):
            n = rng.standard_normal(3); n /= np.linalg.norm(n)
            normal[yy, xx] = orc.pack_normal16i(n)
    f1, n1, d1 = orc.rsm_downsample(flux.view(np.uint16), normal, depth.view(np.uint16))
    f32 = flux.astype(np.float32)
    # flux: SUM of the four texels (energy preserving, downsamplersm.frag:17-22), textureGather order
    s = ((f32[1::2, 0::2] + f32[1::2, 1::2]) + f32[0::2, 1::2]) + f32[0::2, 0::2]
    assert np.array_equal(f1.view(np.float16)[..., :3], s.astype(np.float16)[..., :3])
    dd = depth.astype(np.float32)
    mean = (dd[0::2, 0::2] * 0.5 + dd[0::2, 1::2] * 0.5) * 0.5 + (dd[1::2, 0::2] * 0.5 + dd[1::2, 1::2] * 0.5) * 0.5
    assert np.array_equal(d1.view(np.float16), mean.astype(np.float16))  # MEAN of depth and depth^2 (:32)
    un = np.zeros((r, r, 3))
    for yy in range(r):
        for xx in range(r):
            un[yy, xx] = orc.unpack_normal16i(*normal[yy, xx])
    m = un[0::2, 0::2] + un[0::2, 1::2] + un[1::2, 0::2] + un[1::2, 1::2]
    m /= np.linalg.norm(m, axis=-1, keepdims=True)
    for yy in range(r // 2):
        for xx in range(r // 2):
            assert np.abs(orc.unpack_normal16i(*n1[yy, xx]) - m[yy, xx]).max() < 5e-4


# ---------------------------------------------------------------------------------- stage 5
def test_apply_constant_ambient_known_answer():
    """Every cache holds only SH00 = c: irradiance = ShCosLobeFactor0 * c at every corner, the trilinear weights
    sum to 1, so the pixel is g0 * c * albedo / pi wherever its 8 corner caches exist (cacheApply.frag:59-114)."""
    wl = workloads.cornell(width=96, height=96, rsm_res=16).build()
    o = OracleFrame(wl).allocate()
    c = 0.37
    o.entries[:o.count, [7, 11, 15]] = c
    img = o.apply()
    shaded = img[..., 3] > 0
    assert shaded.sum() > 1000 and np.array_equal(shaded, wl.depth >= 1e-5)
    lut = np.array([orc.srgb8_to_linear(v) for v in range(256)], np.float32)
    albedo = lut[wl.diffuse[..., :3]]
    want = wl.constant.ShCosLobeFactor0 * c * albedo / math.pi
    ratio = img[..., :3][shaded] / want[shaded]
    assert ratio.max() < 1.0 + 1e-5
    assert np.mean(np.abs(ratio - 1.0) < 1e-5) > 0.97  # the rest touch a never-allocated corner
    # negative lobes clamp at zero (lightcache.glsl:178)
    o.entries[:o.count, [7, 11, 15]] = -c
    assert not o.apply()[..., :3].any()


def test_sh_reconstruction_vs_bruteforce_point_sum_sanity():
    """irradiance reconstructed from the SH2 cache ~ direct per-point VPL sum with the cosine
    (bruteforcersm.frag:42-82), up to SH truncation — a sanity bound, not a parity gate."""
    pos, vpls = workloads.sweep(64, 4096)
    cb = _constant()
    e = np.zeros((64, 32), np.float32); e[:, :3] = pos[:, :3]
    orc.light_caches(cb, abi.VolumeInfo(), [_light(4096)], [vpls], None, None, e, 0, 64, 2, False)
    n = np.array([0.0, 1.0, 0.0])
    g0, g1 = cb.ShCosLobeFactor0, cb.ShCosLobeFactor1
    g2, g20, g22 = cb.ShCosLobeFactor2n2_p1_n1, cb.ShCosLobeFactor20, cb.ShCosLobeFactor2p2
    rec = (g0 * e[:, 7] - g1 * n[1] * e[:, 4] + g1 * n[2] * e[:, 8] - g1 * n[0] * e[:, 12]
           - g2 * n[0] * n[1] * e[:, 16] + g2 * n[1] * n[2] * e[:, 20] + g20 * (3 * n[2] ** 2 - 1) * e[:, 19]
           + g2 * n[0] * n[2] * e[:, 24] + g22 * (n[0] ** 2 - n[1] ** 2) * e[:, 28])
    P = pos[:, None, :3].astype(np.float64)
    t = vpls["Position"][None] - P
    d2 = (t * t).sum(-1)
    t /= np.sqrt(d2)[..., None]
    s = np.clip(-(vpls["Normal"][None] * t).sum(-1), 0, 1) / (d2 + vpls["DiscArea"][None])
    direct = (vpls["Flux"][None, :, 0] * s * np.clip(t[..., 1], 0, 1)).sum(1)
    # the uploaded band-2 lobe factor is negative (renderer.cpp:298): only a loose agreement holds
    assert np.corrcoef(rec, direct)[0, 1] > 0.9


# ---------------------------------------------------------------------------------- rows next to the path
def test_fill_rsm_closed_form():
    """fillrsm.frag:32-61: a fragment on the light's axis at distance d with unit base colour gets
    flux = I * 2 (1 - cosHalf) / R^2 (falloff 1, cos 1, the 1/pi 'preponed'), depthLinSq = (d, d^2); off-axis
    values follow the float64 restatement; texels without a fragment keep the clear value."""
    wl = workloads.cornell(rsm_res=16, read_lod=0).build()
    L = wl.spot_lights[0]
    R = 16
    lp = np.array(L.LightPosition[:3], np.float64)
    ld = np.array(L.LightDirection[:3], np.float64)
    rng = np.random.default_rng(3)
    d = rng.uniform(0.5, 5.0, size=(R, R, 1))
    dirs = ld + rng.normal(size=(R, R, 3)) * 0.35
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    dirs[0, 0] = ld
    d[0, 0] = 2.0
    pos = (lp + dirs * d).astype(np.float32)
    nrm = np.tile(np.array([0.0, 0.0, 3.0], np.float32), (R, R, 1))
    base = rng.uniform(0.1, 1.0, size=(R, R, 3)).astype(np.float32)
    base[0, 0] = 1.0
    cov = np.ones((R, R), np.uint8); cov[5, 7] = 0
    fo, no, do = orc.fill_rsm(L, pos, nrm, base, cov)
    flux = fo.view(np.float16).astype(np.float64)[..., :3]
    depth = do.view(np.float16).astype(np.float64)
    I = np.array(L.LightIntensity[:3], np.float64)
    ch = float(L.LightCosHalfAngle)
    assert np.allclose(flux[0, 0], I * 2.0 * (1.0 - ch) / R ** 2, rtol=1e-3)
    assert np.allclose(depth[0, 0], [2.0, 4.0], rtol=1e-3)
    to_light = lp - pos.astype(np.float64)
    dist = np.linalg.norm(to_light, axis=-1)
    cos = np.clip((-to_light / dist[..., None] * ld).sum(-1), 0, 1)
    k = np.clip(cos - ch, 0, 1) / (1 - ch) * (2 * math.pi * (1 - ch) * cos / R / R) / math.pi
    want = base.astype(np.float64) * I * k[..., None]
    want[5, 7] = 0
    assert np.allclose(flux, want, rtol=2e-3, atol=1e-7)
    assert (flux == 0).all(-1).sum() > 1  # outside the cone: falloff 0
    assert not fo[5, 7].any() and not no[5, 7].any() and not do[5, 7].any()
    covered = cov.astype(bool)
    assert np.all(no[..., 0][covered] == 0) and np.all(no[..., 1][covered] == 32767)  # normalize((0,0,3)) = +z


def test_cone_trace_ao_empty_and_full_volume():
    """ambientocclusion.frag:25-89: an empty volume leaves AO = 1; a full one stops every cone at its first sample
    with weight 1, so the occlusion is sum(w) / 6 = (pi/4 + 5 * 3 pi/20) / 6 = pi / 6."""
    wl = workloads.cornell(width=64, height=64, rsm_res=16, read_lod=0, indirect_shadow=True, voxel_resolution=32).build()
    res = 32
    from dynamicradiancevolume_b200 import abi as _abi
    empty = np.zeros(_abi.voxel_chain_bytes(res), np.uint8)
    full = np.full(_abi.voxel_chain_bytes(res), 255, np.uint8)
    out = orc.cone_trace_ao(wl.per_frame, wl.volume, empty, res, wl.depth, wl.normal, out=np.full((64, 64), -1.0, np.float32))
    shaded = wl.depth >= 1e-6
    assert shaded.any() and np.all(out[shaded] == 1.0) and np.all(out[~shaded] == -1.0)
    out = orc.cone_trace_ao(wl.per_frame, wl.volume, full, res, wl.depth, wl.normal, out=np.full((64, 64), -1.0, np.float32))
    # (a pixel whose start point — 1.6 voxels along the normal — lies outside the volume never enters the loop:
    # saturate(p) == p fails, :74, and it keeps AO = 1)
    v = out[shaded]
    inside = v != 1.0
    assert inside.mean() > 0.5 and np.allclose(v[inside], 1.0 - math.pi / 6.0, atol=2e-6)


def test_tonemap_drago_known_answers():
    hdr = np.array([[0.0, 1.0, 3.0, 9.0], [0.5, 0.25, 7.0, 0.0]], np.float32)
    got = orc.tonemap(hdr, 1.0, 1.0)
    assert np.array_equal(got[0], [0.0, 1.0, 2.0])
    got = orc.tonemap(hdr, 2.0, math.log2(2.2))
    want = np.log2(hdr[:, :3].astype(np.float64) * 2.0 + 1.0) / math.log2(2.2)
    assert np.allclose(got, want, rtol=1e-6)


def test_cone_trace_ao_against_float64_restatement():
    """ambientocclusion.frag:25-89 restated independently in float64 python (own trilinear / mip sampler per OpenGL
    4.5 8.14, own ONB and march) on a sample of pixels of a voxelised Cornell box."""
    wl = workloads.cornell(width=48, height=48, rsm_res=16, read_lod=0, indirect_shadow=True, voxel_resolution=32).build()
    o = OracleFrame(wl).prepare_inputs()
    res = 32
    got = orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, res, wl.depth, wl.normal)
    levels, off, r = [], 0, res
    while r >= 1:
        levels.append(o.chain[off:off + r ** 3].reshape(r, r, r).astype(np.float64) / 255.0)  # [z, y, x]
        off += r ** 3
        r //= 2

    def tex(level, p):
        a = levels[level]
        n = a.shape[0]
        u = np.asarray(p) * n - 0.5
        i0 = np.floor(u).astype(int)
        f = u - i0
        acc = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    x, y, z = (int(np.clip(i0[k] + d, 0, n - 1)) for k, d in enumerate((dx, dy, dz)))
                    w = (f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])
                    acc += w * a[z, y, x]
        return acc

    def sample(p, lod):
        lod = min(max(lod, 0.0), len(levels) - 1.0) if lod == lod else 0.0
        l0 = int(math.floor(lod))
        t = lod - l0
        v = tex(l0, p)
        return v if t == 0.0 else v * (1 - t) + tex(min(l0 + 1, len(levels) - 1), p) * t

    PI = 3.14159265358979
    dirs = [(0.0, 1.0, 0.0, PI / 4.0), (0.0, 0.5, 0.866025, 3.0 * PI / 20.0), (0.823639, 0.5, 0.267617, 3.0 * PI / 20.0),
            (0.509037, 0.5, -0.700629, 3.0 * PI / 20.0), (-0.509037, 0.5, -0.700629, 3.0 * PI / 20.0),
            (-0.823639, 0.5, 0.267617, 3.0 * PI / 20.0)]
    ivp = np.array(list(wl.per_frame.InverseViewProjection), np.float64).reshape(4, 4)
    vmin = np.array(wl.volume.VolumeWorldMin[:3], np.float64)
    vs = float(wl.volume.VoxelSizeInWorld)
    rng = np.random.default_rng(11)
    ys, xs = np.nonzero(wl.depth >= 1e-6)
    pick = rng.choice(len(ys), 60, replace=False)
    worst, close_count = 0.0, 0
    for y, x in zip(ys[pick], xs[pick]):
        clip = np.array([(x + 0.5) / 48 * 2 - 1, (y + 0.5) / 48 * 2 - 1, wl.depth[y, x], 1.0])
        w4 = ivp @ clip
        wp = w4[:3] / w4[3]
        px, py = int(wl.normal[y, x, 0]), int(wl.normal[y, x, 1])
        a, z = px * PI / 32768.0, py / 32768.0
        n = np.array([math.cos(a) * math.sqrt(1 - z * z), math.sin(a) * math.sqrt(1 - z * z), z])
        n /= np.linalg.norm(n)
        U = np.cross(n, [0.0, 1.0, 0.0])
        if np.all(np.abs(U) < 1e-4):
            U = np.cross(n, [1.0, 0.0, 0.0])
        U /= np.linalg.norm(U)
        V = np.cross(n, U)
        start = (wp + n * vs * 1.6 - vmin) / (vs * res)
        total = 0.0
        for sx, sy, sz, wgt in dirs:
            d = (sx * V + sy * n + sz * U) / res
            p, step, dist, cw = start.copy(), 1.0, 0.0, 0.0
            s = 0
            while s < 16 and cw < 0.99 and np.all(np.clip(p, 0, 1) == p):
                p = p + d * step
                dist += step
                rad = dist * 0.5
                cw += (1 - cw) * sample(p, math.log2(rad))
                step = rad * 2.0
                s += 1
            total += cw * wgt / 6.0
        want = min(max(1.0 - total, 0.0), 1.0)
        err = abs(want - float(got[y, x]))
        worst = max(worst, err)
        close_count += err < 1e-5
    assert worst <= 2e-3, worst          # a stop test flipped by float32 vs float64 moves AO by <= 1.3e-3
    assert close_count >= 57, close_count


@pytest.mark.parametrize("order", [1, 2])
def test_apply_against_float64_restatement(order):
    """cacheApply.frag:28-195 + lightcache.glsl:109-183 restated independently in float64 python on a sample of
    pixels of a two-cascade frame with transitions (own unprojection, cascade choice, trilinear weights over the
    eight atlas corners, SH evaluation with the max(0, .) per corner, transition blend)."""
    wl = workloads.atrium(width=160, height=90, rsm_res=32, read_lod=0, sh_order=order, cav_resolution=16,
                          first_cascade=8.0, max_caches=8192).build()
    o = OracleFrame(wl).prepare_inputs()
    img = o.frame()
    cb, vi = wl.constant, wl.volume
    R, C = cb.AddressVolumeResolution, cb.NumAddressVolumeCascades
    atlas, E = o.alloc["atlas"], o.entries.astype(np.float64)
    ivp = np.array(list(wl.per_frame.InverseViewProjection), np.float64).reshape(4, 4)
    lut = np.array([orc.srgb8_to_linear(v) for v in range(256)], np.float64)
    g0, g1 = cb.ShCosLobeFactor0, cb.ShCosLobeFactor1
    g2, g20, g22 = cb.ShCosLobeFactor2n2_p1_n1, cb.ShCosLobeFactor20, cb.ShCosLobeFactor2p2
    PI = 3.14159265358979

    def irradiance(addr, n):
        if addr < 0 or addr >= len(E):
            return np.zeros(3)
        e = E[addr]
        irr = e[[7, 11, 15]] * g0 - e[4:7] * (g1 * n[1]) + e[8:11] * (g1 * n[2]) - e[12:15] * (g1 * n[0])
        if order == 2:
            irr = (irr - e[16:19] * (g2 * n[0] * n[1]) + e[20:23] * (g2 * n[1] * n[2])
                   + e[[19, 23, 27]] * (g20 * (3 * n[2] ** 2 - 1)) + e[24:27] * (g2 * n[0] * n[2])
                   + e[28:31] * (g22 * (n[0] ** 2 - n[1] ** 2)))
        return np.maximum(irr, 0.0)

    def lighting(wp, n, c, albedo):
        k = vi.AddressVolumeCascades[c]
        a = (wp - np.array(k.Min[:3], np.float64)) / k.WorldVoxelSize
        b = np.trunc(a).astype(int)
        f = a - b
        total = np.zeros(3)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    x, y, z = b[0] + dx + R * c, b[1] + dy, b[2] + dz
                    addr = -1
                    if 0 <= x < R * C and 0 <= y < R and 0 <= z < R:
                        addr = int(atlas[z, y, x]) - 1
                    w = (f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])
                    total += irradiance(addr, n) * w
        return total * albedo / PI

    rng = np.random.default_rng(21)
    ys, xs = np.nonzero(wl.depth >= 1e-5)
    pick = rng.choice(len(ys), 160, replace=False)
    good = 0
    saw_transition = False
    for y, x in zip(ys[pick], xs[pick]):
        clip = np.array([(x + 0.5) / wl.width * 2 - 1, (y + 0.5) / wl.height * 2 - 1, wl.depth[y, x], 1.0])
        w4 = ivp @ clip
        wp = w4[:3] / w4[3]
        c = C - 1
        for ci in range(C - 1):
            k = vi.AddressVolumeCascades[ci]
            if np.all(wp <= np.array(k.DecisionMax[:3])) and np.all(wp >= np.array(k.DecisionMin[:3])):
                c = ci
                break
        a, z = int(wl.normal[y, x, 0]) * PI / 32768.0, int(wl.normal[y, x, 1]) / 32768.0
        n = np.array([math.cos(a) * math.sqrt(1 - z * z), math.sin(a) * math.sqrt(1 - z * z), z])
        n /= np.linalg.norm(n)
        albedo = lut[wl.diffuse[y, x, :3]]
        col = lighting(wp, n, c, albedo)
        if wl.transitions and c < C - 1:
            k = vi.AddressVolumeCascades[c]
            md = min((np.array(k.DecisionMax[:3]) - wp).min(), (wp - np.array(k.DecisionMin[:3])).min())
            tr = min(max(1.0 - md / (k.WorldVoxelSize * vi.CAVTransitionZoneSize), 0.0), 1.0)
            if tr > 0.0:
                saw_transition = True
                col = col * (1 - tr) + lighting(wp, n, c + 1, albedo) * tr
        ref = img[y, x, :3].astype(np.float64)
        if np.all(np.abs(col - ref) <= 1e-7 + 2e-5 * np.maximum(np.abs(col), np.abs(ref))):
            good += 1
    # the few that differ sit on a float32 cell / cascade boundary where a neighbouring corner cache was never
    # allocated (the trilinear blend is continuous only where all eight corners exist)
    assert good >= 155, good
    assert img[..., :3].max() > 1e-3


def _morton(k):
    x = y = 0
    for b in range(16):
        x |= ((k >> (2 * b)) & 1) << b
        y |= ((k >> (2 * b + 1)) & 1) << b
    return x, y


def test_vpl_list_and_shadow_blocks_against_float64_restatement():
    """cacheLightingRSM.comp:137-163 (VPL load) and :168-192 (the cache-independent half of the indirect-shadow
    sample) restated independently in float64 python: Morton order, unprojection along the light ray, disc area,
    bilinear clamp-to-edge fetch of depthLinSq at the shadow LOD, distToSphereRad."""
    wl = workloads.atrium(width=64, height=64, rsm_res=128, read_lod=1, sh_order=1, indirect_shadow=True,
                          voxel_resolution=32, shadow_lod=2).build()
    o = OracleFrame(wl).prepare_inputs()
    L = wl.spot_lights[0]
    R = int(L.RSMReadResolution)
    read_level = workloads.rsm_read_level(L)
    flux, normal, depth = (np.asarray(a) for a in o.levels[0][read_level])
    fl = flux.view(np.float16).astype(np.float64)
    dp = depth.view(np.float16).astype(np.float64)
    ilvp = np.array(list(L.InverseLightViewProjection), np.float64).reshape(4, 4)
    lp = np.array(L.LightPosition[:3], np.float64)
    PI = 3.14159265358979

    def ray_point(u, v, d):
        w = ilvp @ np.array([u * 2 - 1, v * 2 - 1, 0.0, 1.0])
        dirv = w[:3] / w[3] - lp
        return lp + dirv / np.linalg.norm(dirv) * d

    vpls = o.vpls[0]
    rng = np.random.default_rng(9)
    for k in rng.choice(R * R, 300, replace=False):
        x, y = _morton(int(k))
        d = dp[y, x, 0]
        want = ray_point((x + 0.5) / R, (y + 0.5) / R, d)
        assert np.allclose(vpls["Position"][k], want, rtol=1e-5, atol=1e-5)
        assert np.isclose(vpls["DiscArea"][k], d * d * L.ValAreaFactor, rtol=1e-5)
        assert np.array_equal(vpls["Flux"][k], fl[y, x, :3].astype(np.float32))
        a, z = int(normal[y, x, 0]) * PI / 32768.0, int(normal[y, x, 1]) / 32768.0
        n = np.array([math.cos(a) * math.sqrt(1 - z * z), math.sin(a) * math.sqrt(1 - z * z), z])
        assert np.allclose(vpls["Normal"][k], n / np.linalg.norm(n), atol=2e-6)
    # shadow blocks
    lod = int(L.IndirectShadowComputationLod)
    interval = int(L.IndirectShadowComputationSampleInterval)
    assert interval == 4 ** lod
    dl = np.asarray(o.levels[0][read_level + lod][2]).view(np.float16).astype(np.float64)  # [Rl, Rl, 2]
    Rl = R >> lod
    blocks = o.blocks[0]
    assert len(blocks) == R * R // interval
    for b in rng.choice(len(blocks), 120, replace=False):
        x, y = _morton(int(b) * interval)
        u, v = (x + L.IndirectShadowSamplingOffset) / R, (y + L.IndirectShadowSamplingOffset) / R
        fx, fy = u * Rl - 0.5, v * Rl - 0.5
        x0, y0 = math.floor(fx), math.floor(fy)
        tx, ty = fx - x0, fy - y0
        cl = lambda i: min(max(i, 0), Rl - 1)
        s = lambda c: ((dl[cl(y0), cl(x0), c] * (1 - tx) + dl[cl(y0), cl(x0 + 1), c] * tx) * (1 - ty)
                       + (dl[cl(y0 + 1), cl(x0), c] * (1 - tx) + dl[cl(y0 + 1), cl(x0 + 1), c] * tx) * ty)
        m1, m2 = s(0), s(1)
        want_pos = ray_point(u, v, m1)
        var = max(m2 - m1 * m1, 0.0)
        want_k = max(L.IndirectShadowComputationSuperValWidth, math.sqrt(var) * 2.0 / m1) if m1 > 0 else None
        assert np.allclose(blocks["AverageValPos"][b], want_pos, rtol=2e-5, atol=2e-5)
        if want_k is not None and var > 1e-4 * m1 * m1:  # away from the cancellation-dominated variance ~ 0 case
            assert np.isclose(blocks["DistToSphereRad"][b], want_k, rtol=2e-3)
        elif want_k is not None:
            assert blocks["DistToSphereRad"][b] >= L.IndirectShadowComputationSuperValWidth * (1 - 1e-6)


def _py_sampler(chain, res):
    """An independent float64 sampler3D (linear / mip-linear / clamp to edge, OpenGL 4.5 8.14) over a voxel chain."""
    levels, off, r = [], 0, res
    while r >= 1:
        levels.append(chain[off:off + r ** 3].reshape(r, r, r).astype(np.float64) / 255.0)  # [z, y, x]
        off += r ** 3
        r //= 2

    def tex(level, p):
        a = levels[level]
        n = a.shape[0]
        u = np.asarray(p, np.float64) * n - 0.5
        i0 = np.floor(u).astype(int)
        f = u - i0
        acc = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    x, y, z = (int(np.clip(i0[k] + d, 0, n - 1)) for k, d in enumerate((dx, dy, dz)))
                    acc += ((f[0] if dx else 1 - f[0]) * (f[1] if dy else 1 - f[1]) * (f[2] if dz else 1 - f[2])) * a[z, y, x]
        return acc

    def sample(p, lod):
        lod = min(max(lod, 0.0), len(levels) - 1.0) if lod == lod else 0.0
        l0 = int(math.floor(lod))
        t = lod - l0
        v = tex(l0, p)
        return v if t == 0.0 else v * (1 - t) + tex(min(l0 + 1, len(levels) - 1), p) * t
    return sample


def test_shadow_cone_against_float64_restatement():
    """cacheLightingRSM.comp:195-230 restated independently in float64 python (own sampler, own march) for random
    cache / block pairs through a voxelised Cornell box."""
    wl = workloads.cornell(width=32, height=32, rsm_res=16, read_lod=0, indirect_shadow=True, voxel_resolution=32).build()
    o = OracleFrame(wl).prepare_inputs().allocate()
    res = 32
    sample = _py_sampler(o.chain, res)
    vi = wl.volume
    vmin, vs = np.array(vi.VolumeWorldMin[:3], np.float64), float(vi.VoxelSizeInWorld)
    rng = np.random.default_rng(4)
    worst, agree, shadowed = 0.0, 0, 0
    for _ in range(120):
        pos = o.entries[rng.integers(o.count), :3].astype(np.float64)
        blk = o.blocks[0][rng.integers(len(o.blocks[0]))]
        kk = float(blk["DistToSphereRad"])
        to = blk["AverageValPos"].astype(np.float64) - pos
        light_dist = np.linalg.norm(to)
        d = to / light_dist / res
        p = (pos - vmin) / (vs * res) + d * 2.0
        occ, step, dist = 0.0, 1.0, 0.0
        goal = light_dist / vs - 2.0
        r2s = 2.0 / (1.0 - kk)
        for s in range(32):
            p = p + d * step
            dist += step
            rad = dist * kk
            occ += (1 - occ) * sample(p, math.log2(rad))
            if dist >= goal:
                break
            step = max(1.0, rad * r2s)
        want = min(max(1.0 - occ, 0.0), 1.0)
        got = orc.cone_trace(vi, o.chain, res, pos, np.array([blk]))
        err = abs(want - got)
        worst = max(worst, err)
        agree += err < 1e-5
        shadowed += want < 0.99
    assert worst < 2e-2, worst        # a float32 / float64 difference in `dist >= goal` adds or drops one far sample
    assert agree >= 114, agree
    assert shadowed > 10


def test_voxelizer_is_conservative_and_tight():
    """voxelize.{vert,geom,frag} as a coverage set against an exact triangle / box separating-axis test in float64,
    for random triangles: every voxel whose core (a cube of 0.4 voxels) the triangle crosses is set and at least
    97 % of all voxels it touches at all (the scheme is conservative in the raster plane; in depth it extrapolates
    the ORIGINAL vertex depths over the dilated triangle and widens by 1.414 |grad z|, voxelize.geom:83-88 /
    voxelize.frag:39-57, so a triangle grazing a voxel corner can be missed — in the reference too); nothing farther
    than two voxels from the triangle is set."""
    res = 16
    vi = _volume(res)
    rng = np.random.default_rng(12)
    centres = (np.stack(np.meshgrid(np.arange(res), np.arange(res), np.arange(res), indexing="ij"), -1) + 0.5).reshape(-1, 3)  # x, y, z

    def tri_box_overlap(tri, c, h):
        v = tri - c
        e = [v[1] - v[0], v[2] - v[1], v[0] - v[2]]
        axes = [np.eye(3)[i] for i in range(3)] + [np.cross(e[0], e[1])] + [np.cross(np.eye(3)[i], ej) for i in range(3) for ej in e]
        for a in axes:
            if not np.any(a):
                continue
            pr = v @ a
            rad = h * np.abs(a).sum()
            if pr.min() > rad or pr.max() < -rad:
                return False
        return True

    n_touched = n_missed = 0
    for _ in range(12):
        tri = rng.uniform(1.5, res - 1.5, size=(3, 3))
        vol = orc.voxelize(vi, res, tri.reshape(1, 9).astype(np.float32)).reshape(res, res, res) > 0  # z, y, x
        lo, hi = np.floor(tri.min(0)).astype(int) - 3, np.ceil(tri.max(0)).astype(int) + 3
        touched = np.zeros((res, res, res), bool)
        core = np.zeros((res, res, res), bool)
        near = np.zeros((res, res, res), bool)
        for c in centres:
            if np.any(c < lo) or np.any(c > hi):
                continue
            x, y, z = (int(v) for v in c)
            if tri_box_overlap(tri, c, 0.5 - 1e-9):
                touched[z, y, x] = True
                core[z, y, x] = tri_box_overlap(tri, c, 0.2)
            if tri_box_overlap(tri, c, 2.5):
                near[z, y, x] = True
        assert core.any()
        assert not (core & ~vol).any(), "a voxel whose core the triangle crosses is not set"
        assert not (vol & ~near).any(), "a voxel farther than two voxels from the triangle is set"
        n_touched += int(touched.sum())
        n_missed += int((touched & ~vol).sum())
    assert n_missed <= 0.03 * n_touched, (n_missed, n_touched)


@pytest.mark.parametrize("name,make", [
    ("cornell", lambda: workloads.cornell(width=96, height=80, rsm_res=16)),
    ("atrium-transitions", lambda: workloads.atrium(width=160, height=90, rsm_res=32, read_lod=0, cav_resolution=16,
                                                    first_cascade=8.0, max_caches=8192)),
    ("atrium-3casc-ragged", lambda: workloads.atrium(width=150, height=70, rsm_res=32, read_lod=0, cascades=3,
                                                     cav_resolution=16, first_cascade=4.0, max_caches=8192)),
])
def test_allocation_set_against_vectorised_float32_restatement(name, make):
    """cacheGather.comp:93-164 restated independently as whole-image numpy float32 array code (every * and +
    rounded separately, like the oracle's policy): world positions, cascade choice, cell ids, the two neighbour
    dedupe predicates on the shader's 16x16 tiles, the eight corner cells. The allocated cell SET must be identical."""
    wl = make().build()
    cb, vi = wl.constant, wl.volume
    W, H, R, C = cb.BackbufferResolution[0], cb.BackbufferResolution[1], cb.AddressVolumeResolution, cb.NumAddressVolumeCascades
    f = np.float32
    d = wl.depth.astype(f)
    xs, ys = np.meshgrid(np.arange(W, dtype=f), np.arange(H, dtype=f))
    sx = (xs + f(0.5)) / f(W) * f(2) - f(1)
    sy = (ys + f(0.5)) / f(H) * f(2) - f(1)
    m = np.array(list(wl.per_frame.InverseViewProjection), f).reshape(4, 4)
    one = np.ones_like(d)
    row = lambda r: ((m[r, 0] * sx + m[r, 1] * sy) + m[r, 2] * d) + m[r, 3] * one
    w = row(3)
    with np.errstate(divide="ignore", invalid="ignore"):
        wp = np.stack([row(0) / w, row(1) / w, row(2) / w], -1)
    valid = d > f(0.0001)

    def inside(c):
        k = vi.AddressVolumeCascades[c]
        return np.all(wp <= np.array(k.DecisionMax[:3], f), -1) & np.all(wp >= np.array(k.DecisionMin[:3], f), -1)

    casc = np.full((H, W), C - 1, int)
    for c in range(C - 2, -1, -1):
        casc = np.where(inside(c), c, casc)

    def cell(c_arr):
        out = np.zeros((H, W), np.int64)
        for c in range(C):
            k = vi.AddressVolumeCascades[c]
            with np.errstate(invalid="ignore"):
                g = (wp - np.array(k.Min[:3], f)) / f(k.WorldVoxelSize)
            gi = np.clip(np.trunc(np.nan_to_num(g, nan=0.0, posinf=1e9, neginf=-1e9)), 0, R - 1).astype(np.int64)
            idc = gi[..., 0] + gi[..., 1] * R + gi[..., 2] * R * R + c * R ** 3
            out = np.where(c_arr == c, idc, out)
        return out

    T1 = np.where(valid, cell(casc), -1)
    T2 = np.full((H, W), -1, np.int64)
    if wl.transitions:
        tr = np.zeros((H, W), f)
        for c in range(C - 1):
            k = vi.AddressVolumeCascades[c]
            with np.errstate(invalid="ignore"):
                md = np.minimum((np.array(k.DecisionMax[:3], f) - wp).min(-1), (wp - np.array(k.DecisionMin[:3], f)).min(-1))
                t = np.clip(f(1) - md / (f(k.WorldVoxelSize) * f(vi.CAVTransitionZoneSize)), 0, 1)
            tr = np.where(casc == c, t, tr)
        second = valid & (tr > 0) & (casc < C - 1)
        T2 = np.where(second, cell(np.minimum(casc + 1, C - 1)), -1)

    # pad to whole tiles with -1 (threads outside the image leave 0xFFFFFFFF in shared memory)
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    pad = lambda a: np.pad(a, ((0, Hp - H), (0, Wp - W)), constant_values=-1)
    T1p, T2p, cp = pad(T1), pad(T2), pad(casc)
    ly, lx = np.meshgrid(np.arange(Hp) % 16, np.arange(Wp) % 16, indexing="ij")
    yy, xx = np.meshgrid(np.arange(Hp), np.arange(Wp), indexing="ij")
    up, left = np.where(ly > 0, yy - 1, yy), np.where(lx > 0, xx - 1, xx)
    trig1 = (((T1p[up, xx] != T1p) & (T1p[yy, left] != T1p) & (T1p[up, left] != T1p)) | ((lx == 0) & (ly == 0))) & (T1p != -1)
    dn, right = np.where(ly < 15, yy + 1, yy), np.where(lx < 15, xx + 1, xx)
    # `lookUpThread == ivec2(LOCAL_SIZE-1)` holds for the clamped look-up, i.e. for local x and y >= 14 (:157)
    trig2 = (((T2p[dn, xx] != T2p) & (T2p[yy, right] != T2p) & (T2p[dn, right] != T2p)) | ((lx >= 14) & (ly >= 14))) & (T2p != -1)
    ids = set()
    for trig, T, cc in ((trig1, T1p, cp), (trig2, T2p, np.minimum(cp + 1, C - 1))):
        for coord, c in zip(T[trig], cc[trig]):
            local = int(coord) - R ** 3 * int(c)
            bz, by, bx = local // (R * R), (local // R) % R, local % R
            for ox in (0, 1):
                for oy in (0, 1):
                    for oz in (0, 1):
                        if bx + ox < R and by + oy < R and bz + oz < R:
                            ids.add((bx + ox) + (by + oy) * R + (bz + oz) * R * R + int(c) * R ** 3)
    got = orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.depth)
    assert len(got) > 50
    assert np.array_equal(np.array(sorted(ids), np.int32), got)
