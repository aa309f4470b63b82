"""Pins the oracle to the REFERENCE'S OWN SOURCE TEXT.

oracle/_ref/libdrv_ref.so is the reference's unmodified GLSL (shader/cacheGather.comp, cachePrepareLighting.comp,
cacheLightingRSM.comp, cacheApply.frag with lightcache.glsl / utils.glsl / globalubos.glsl, voxelblend.comp,
voxelmipmap.comp, downsamplersm.frag), rewritten mechanically into C++ by oracle/ref/glsl2cpp.py and compiled with g++
against oracle/ref/glsl_compat.h (work groups as fibers, shared memory, barriers, atomics, software samplers).
These tests run the hand-written restatement (oracle/*.cpp) and that library on the same inputs — BASELINE configs[0]
at full size, reduced configs[1] and configs[2] — and require BIT-FOR-BIT equality: allocated cell set and entry
positions, counter / indirect arguments, the VPL list the shader derives in shared memory, every SH coefficient
(SH1 / SH2, with and without cone-traced shadows, two lights), the applied image (with and without cascade
transitions), the voxel blend, the RSM mip rule; the voxel mip chain is exact except on rounding ties of the
UNORM8 store, which GL leaves open (counted and bounded to one code).

No GPU is needed. The library is built here from /root/reference (`make -C oracle ref`); on a box without the
reference checkout the prebuilt file from the snapshot is used, and without either the tests skip.
"""
import numpy as np
import pytest

import workloads
from oracle import binding as orc
from oracle.frame import OracleFrame
from oracle.ref import binding as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="neither oracle/_ref/libdrv_ref.so nor /root/reference is present")


@pytest.fixture(scope="module", autouse=True)
def _built():
    ref.build()
    ref.load()


def _cells(atlas):
    return np.flatnonzero(atlas.ravel())


def _by_cell(atlas, entries):
    """entries re-ordered by linear atlas position (the reference's indices depend on execution order)."""
    flat = atlas.ravel()
    idx = np.flatnonzero(flat)
    return idx, entries[flat[idx].astype(np.int64) - 1]


WORKLOADS = {
    # BASELINE configs[0] at full size
    "c1": lambda: workloads.cornell(),
    # configs[1] reduced: 2 cascades + transitions, 1024^2 RSM read at LOD 4 (64^2 = 4096 VPLs)
    "c2r": lambda: workloads.atrium(width=480, height=270, rsm_res=1024, read_lod=4, cav_resolution=32),
    # three cascades, no transitions, SH2
    "c2r3": lambda: workloads.atrium(width=320, height=180, rsm_res=256, read_lod=2, cav_resolution=16, cascades=3,
                                     first_cascade=4.0, transition=0.0, sh_order=2),
    # configs[2] reduced: cone-traced shadows through a 64^3 voxel chain, SH2, shadow LOD 1
    "c3r": lambda: workloads.atrium(width=320, height=180, rsm_res=256, read_lod=2, cav_resolution=32, sh_order=2,
                                    indirect_shadow=True, voxel_resolution=64, shadow_lod=1),
    # cornell with shadows, SH1, default shadow LOD 2, two lights
    "c1s": lambda: workloads.cornell(width=256, height=256, rsm_res=128, read_lod=1, indirect_shadow=True,
                                     voxel_resolution=32),
}
_cache = {}


def _frame(name):
    if name not in _cache:
        wl = WORKLOADS[name]().build()
        _cache[name] = OracleFrame(wl).prepare_inputs().allocate()
    return _cache[name]


@pytest.mark.parametrize("name", ["c1", "c2r", "c2r3", "c3r"])
def test_allocation_matches_the_reference_shader(name):
    o = _frame(name)
    wl = o.wl
    r = ref.allocate_caches(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.depth, wl.sh_order)
    assert r["count"] == o.count > 0
    # the allocated SET
    assert np.array_equal(_cells(r["atlas"]), _cells(o.alloc["atlas"]))
    # every cell's entry: Position bit-exact, SH zeroed (cacheGather.comp:65-83)
    ci, er = _by_cell(r["atlas"], r["entries"])
    _, eo = _by_cell(o.alloc["atlas"], o.alloc["entries"])
    assert np.array_equal(er.view(np.uint32), eo.view(np.uint32))
    # indices are a permutation of 0..count-1; cachePrepareLighting.comp:8-14
    assert np.array_equal(np.sort(r["atlas"].ravel()[ci]), np.arange(1, o.count + 1, dtype=np.uint32))
    c = r["counter"]
    assert (c.NumCacheLightingThreadGroupsX, c.NumCacheLightingThreadGroupsY, c.NumCacheLightingThreadGroupsZ,
            c.TotalLightCacheCount) == ((o.count + 63) // 64, 1, 1, o.count)
    oc = o.alloc["counter"]
    assert (oc.NumCacheLightingThreadGroupsX, oc.TotalLightCacheCount) == (c.NumCacheLightingThreadGroupsX, o.count)


def _light_both(o, count, tap):
    wl = o.wl
    eo = o.alloc["entries"].copy()
    o.light(first=0, count=count, entries=eo)
    er = o.alloc["entries"].copy()
    reads = [workloads.rsm_read_level(s) for s in wl.spot_lights]
    taps = ref.light_caches(wl.constant, wl.per_frame, wl.volume, wl.spot_lights, o.levels, reads, o.chain,
                            wl.voxel_resolution, er, count, wl.sh_order, wl.indirect_shadow, tap=tap)
    return eo, er, taps


@pytest.mark.parametrize("name", ["c1", "c2r", "c2r3"])
def test_vpl_list_and_unshadowed_sh_match_the_reference_shader(name):
    o = _frame(name)
    eo, er, taps = _light_both(o, o.count, tap=True)
    # the VPL list as the shader derives it into shared memory (cacheLightingRSM.comp:135-163)
    for v, t in zip(o.vpls, taps):
        assert np.array_equal(t[:, 0:3].view(np.uint32), v["Flux"].view(np.uint32))
        assert np.array_equal(t[:, 3].view(np.uint32), v["DiscArea"].view(np.uint32))
        assert np.array_equal(t[:, 4:7].view(np.uint32), v["Position"].view(np.uint32))
        assert np.array_equal(t[:, 7:10].view(np.uint32), v["Normal"].view(np.uint32))
    n = o.count
    assert np.abs(eo[:n, 4:]).max() > 0
    assert np.array_equal(er[:n].view(np.uint32), eo[:n].view(np.uint32)), "SH coefficients differ from the reference shader"


@pytest.mark.parametrize("name,count", [("c3r", 512), ("c1s", 320)])
def test_cone_traced_sh_matches_the_reference_shader(name, count):
    """INDIRECT_SHADOW (cacheLightingRSM.comp:167-232) on the first `count` entries (whole 64-entry groups)."""
    o = _frame(name)
    count = min(count, o.count // 64 * 64)
    eo, er, _ = _light_both(o, count, tap=False)
    assert np.abs(eo[:count, 4:]).max() > 0
    assert np.array_equal(er[:count].view(np.uint32), eo[:count].view(np.uint32))
    # shadows bite: the unshadowed result differs
    wl = o.wl
    e2 = o.alloc["entries"].copy()
    orc.light_caches(wl.constant, wl.volume, wl.spot_lights, o.vpls, [None] * len(o.vpls), None, e2, 0, count,
                     wl.sh_order, False)
    assert not np.array_equal(e2[:count], eo[:count])


def test_two_lights_accumulate_like_the_reference_shader():
    wl = workloads.atrium(width=160, height=90, rsm_res=128, read_lod=1, cav_resolution=16, num_lights=2).build()
    o = OracleFrame(wl).prepare_inputs().allocate()
    eo, er, _ = _light_both(o, o.count, tap=False)
    assert np.array_equal(er[:o.count].view(np.uint32), eo[:o.count].view(np.uint32))


@pytest.mark.parametrize("name", ["c1", "c2r", "c2r3", "c3r"])
def test_apply_matches_the_reference_shader(name):
    o = _frame(name)
    wl = o.wl
    n = min(o.count, 2048) if wl.indirect_shadow else o.count
    e = o.alloc["entries"].copy()
    o.light(first=0, count=n, entries=e)
    img_o = o.apply(entries=e)
    img_r = ref.apply_caches(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal,
                             wl.diffuse, o.alloc["atlas"], e)
    assert img_o[..., :3].max() > 0
    assert np.array_equal(img_r.view(np.uint32), img_o.view(np.uint32)), "applied image differs from cacheApply.frag"


def test_voxel_blend_and_mips_match_the_reference_shaders():
    o = _frame("c3r")
    res = o.wl.voxel_resolution
    rng = np.random.default_rng(7)
    old = rng.integers(0, 256, res ** 3, dtype=np.uint8)
    for k in (1, 9, 255):
        a, b = old.copy(), old.copy()
        orc.voxel_blend(a, o.target, res, k / 255.0)
        ref.voxel_blend(b, o.target, res, k / 255.0)
        assert np.array_equal(a, b), "voxelblend.comp, adaption %d/255" % k
    for level0 in (o.chain[: res ** 3], old):
        co = orc.voxel_chain(level0, res)
        off, r, ties = 0, res, 0
        while r > 1:
            # one mip step of voxelmipmap.comp on the oracle's level (a tie decided differently would otherwise
            # propagate into the coarser levels)
            h = r // 2
            src = co[off: off + r ** 3]
            step_r = ref.voxel_chain(src, r)[r ** 3: r ** 3 + h ** 3]
            step_o = co[off + r ** 3: off + r ** 3 + h ** 3]
            diff = np.flatnonzero(step_o != step_r)
            s3 = src.reshape(r, r, r).astype(np.int32)
            sums = sum(s3[dz::2, dy::2, dx::2] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)).ravel()
            # UNORM8 store of a value exactly between two codes: GL leaves the tie open; everything else is exact
            assert np.all((sums[diff] % 8) == 4), "a non-tie voxel differs from voxelmipmap.comp"
            assert np.all(np.abs(step_o[diff].astype(int) - step_r[diff].astype(int)) <= 1)
            ties += int(np.count_nonzero(sums % 8 == 4))
            off, r = off + r ** 3, h
        assert ties > 0


def test_rsm_mip_rule_matches_the_reference_shader():
    o = _frame("c2r")
    lv = o.levels[0]
    for l in range(0, 5):
        fo, no, do = orc.rsm_downsample(*lv[l])
        fr, nr, dr = ref.rsm_downsample(*lv[l])
        assert np.array_equal(fo, fr) and np.array_equal(do, dr), "level %d" % l
        assert np.array_equal(no, nr), "packed normals of level %d" % l


# ---- the rows next to the path ------------------------------------------------------------------
def test_fill_rsm_matches_the_reference_shader():
    """shader/fillrsm.frag: flux and depthLinSq halfs bit-exact; the packed normal within one int16 code (the shader
    normalises the interpolated normal, pushes it through the tangent frame and normalises again)."""
    res = 96
    wl = workloads.cornell(rsm_res=128, read_lod=0).build(render=False)
    light = wl.spot_lights[0]
    rng = np.random.default_rng(5)
    lp, ld = np.array(light.LightPosition[:3], np.float32), np.array(light.LightDirection[:3], np.float32)
    dirs = rng.normal(size=(res, res, 3)).astype(np.float32) * 0.6 + ld
    pos = (lp + dirs * rng.uniform(0.5, 6.0, size=(res, res, 1)).astype(np.float32)).astype(np.float32)
    nrm = (rng.normal(size=(res, res, 3)) * rng.uniform(0.2, 3.0, size=(res, res, 1))).astype(np.float32)
    nrm[0, 0], nrm[0, 1] = (0.0, 2.0, 0.0), (0.0, -1.0, 0.0)
    base = rng.uniform(0.0, 1.0, size=(res, res, 3)).astype(np.float32)
    cov = (rng.uniform(size=(res, res)) > 0.25).astype(np.uint8)
    fo, no, do = orc.fill_rsm(light, pos, nrm, base, cov)
    fr, nr, dr = ref.fill_rsm(light, pos, nrm, base, cov)
    assert fo[..., :3].max() > 0
    assert np.array_equal(fo, fr) and np.array_equal(do, dr)
    dn = np.abs(no.astype(np.int32) - nr.astype(np.int32))
    dn[..., 0] = np.minimum(dn[..., 0], 65536 - dn[..., 0])
    assert dn.max() <= 1
    assert np.count_nonzero(dn) < 0.02 * dn.size


def test_ambient_occlusion_matches_the_reference_shader():
    """shader/ambientocclusion.frag: the oracle folds sin(pi/6) to 0.5 as the GLSL compiler does; the run-time sinf of
    the compiled shader text gives the neighbouring float, and the loop's `coneWeight < 0.99` test can flip on such a
    difference — so this row is compared with the f-row tolerance, and the exact fraction is reported."""
    o = _frame("c3r")
    wl = o.wl
    ao_o = orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal)
    ao_r = ref.cone_trace_ao(wl.constant, wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal)
    assert ao_o.max() - ao_o.min() > 0.2
    d = np.abs(ao_o - ao_r)
    assert d.max() <= 2e-3
    assert np.mean(d <= 1e-6) > 0.99


def test_tonemap_matches_the_reference_shader():
    rng = np.random.default_rng(3)
    hdr = rng.uniform(0.0, 8.0, size=(257, 4)).astype(np.float32)
    a = orc.tonemap(hdr, 1.7, np.float32(np.log2(1.2 + 1.0)))
    b = ref.tonemap(hdr, 1.7, np.float32(np.log2(1.2 + 1.0)))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- indirect specular -------------------------------------------------------------------------
def _rough_metal(wl, seed=11):
    """A synthetic roughness / metallic G-buffer plane (RG8, renderer.cpp:469): smooth bands + noise."""
    rng = np.random.default_rng(seed)
    H, W = wl.depth.shape
    y, x = np.mgrid[0:H, 0:W]
    rough = (40 + 180 * (0.5 + 0.5 * np.sin(x * 0.05)) * (0.5 + 0.5 * np.cos(y * 0.07))).astype(np.uint8)
    metal = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    return np.ascontiguousarray(np.stack([rough, metal], -1))


@pytest.mark.parametrize("name,sh_order,shadow", [("spec_a", 1, False), ("spec_b", 2, True)])
def test_indirect_specular_matches_the_reference_shaders(name, sh_order, shadow):
    """INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE: SH with the flux early-out, the R11F_G11F_B10F environment-map
    atlas (every texel, including the ones neighbouring maps spill into), its mip chain, the hole filling and the
    applied image — bit for bit."""
    import dynamicradiancevolume_b200 as drv
    if shadow:
        wl = workloads.atrium(width=200, height=112, rsm_res=128, read_lod=1, cav_resolution=16, sh_order=2, transition=0.0,
                              indirect_shadow=True, voxel_resolution=32, shadow_lod=1, max_caches=4096).build()
    else:
        wl = workloads.atrium(width=200, height=112, rsm_res=128, read_lod=1, cav_resolution=16, sh_order=1,
                              max_caches=4096).build()
    drv.pack_specular(wl.constant, wl.max_caches, 16)
    o = OracleFrame(wl).prepare_inputs().allocate()
    n = o.count
    assert 64 < n < wl.max_caches
    total_texels, sizes = orc.specular_mip_texels(wl.constant)
    eo = o.alloc["entries"].copy()
    mo = orc.light_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.spot_lights, o.vpls, o.blocks, o.chain, eo, n,
                                   wl.sh_order, wl.indirect_shadow)
    er = o.alloc["entries"].copy()
    reads = [workloads.rsm_read_level(s) for s in wl.spot_lights]
    mr = ref.light_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.spot_lights, o.levels, reads, o.chain,
                                   wl.voxel_resolution, er, n, wl.sh_order, wl.indirect_shadow, total_texels)
    assert np.array_equal(er[:n].view(np.uint32), eo[:n].view(np.uint32)), "SH with the :241 early-out"
    lvl0 = sizes[0] ** 2
    assert np.count_nonzero(mo[:lvl0]) > 1000
    assert np.array_equal(mo[:lvl0], mr[:lvl0]), "environment-map atlas"
    # Renderer::PrepareSpecularEnvmaps: mip chain, then two levels of hole filling
    orc.specular_mips(wl.constant, n, mo)
    ref.specular_mips(wl.constant, n, mr)
    assert np.array_equal(mo, mr), "specularenvmap_mipmap.frag"
    assert np.count_nonzero(mo[lvl0:]) > 100
    orc.specular_fill_holes(wl.constant, n, 2, mo)
    ref.specular_fill_holes(wl.constant, n, 2, mr)
    assert np.array_equal(mo, mr), "specularenvmap_fillholes.frag"
    rm = _rough_metal(wl)
    img_o = orc.apply_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal,
                                      wl.diffuse, rm, o.alloc["atlas"], eo, mo)
    img_r = ref.apply_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal,
                                      wl.diffuse, rm, o.alloc["atlas"], eo, mr)
    assert np.array_equal(img_r.view(np.uint32), img_o.view(np.uint32)), "cacheApply.frag with INDIRECT_SPECULAR"
    # the specular term is there: the diffuse-only image differs
    plain = o.apply(entries=eo)
    assert np.abs(plain[..., :3] - img_o[..., :3]).max() > 1e-4
