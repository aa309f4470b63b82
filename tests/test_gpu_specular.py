"""GPU parity of SURVEY.md 8(f) row f4 — indirect specular through per-cache environment maps (specular.cu) — through
the C-ABI against the CPU oracle (oracle/specular.cpp, itself pinned bit for bit to the reference's shaders with
INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE: tests/test_oracle_vs_ref.py).

The atlas is R11F_G11F_B10F and every contribution is rounded to 6 / 5 mantissa bits as it is added, so this row has
the gates of an f-row, stated here: the SH (with the shader's flux early-out) within the path's 1e-3 / 1e-5 gate; the
environment-map texels of the maps' interiors equal to the oracle's in >= 99 % of the non-zero texels and within a few
codes of each channel's small float in 99.99 % of them (a contribution whose atan / visibility differs in the last bit
can fall into the neighbouring texel or round the other way); the texels neighbouring maps spill into — a data race in
the reference, serialised differently here — within 25 %; the final image within 2e-2 relative + 1e-4 absolute.
"""
import numpy as np
import pytest

import dynamicradiancevolume_b200 as drv
import workloads
from dynamicradiancevolume_b200 import abi
from oracle import binding as orc
from oracle.frame import OracleFrame, close

pytestmark = pytest.mark.gpu


def _rough_metal(wl, seed=11):
    rng = np.random.default_rng(seed)
    H, W = wl.depth.shape
    y, x = np.mgrid[0:H, 0:W]
    rough = (40 + 180 * (0.5 + 0.5 * np.sin(x * 0.05)) * (0.5 + 0.5 * np.cos(y * 0.07))).astype(np.uint8)
    metal = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    return np.ascontiguousarray(np.stack([rough, metal], -1))


def _unpack(t):
    """uint32 R11F_G11F_B10F -> float64 [.., 3] and the per-channel codes."""
    codes = np.stack([t & 0x7FF, (t >> 11) & 0x7FF, t >> 22], -1).astype(np.int64)
    out = np.zeros(codes.shape, np.float64)
    for c, mb in enumerate((6, 6, 5)):
        v = codes[..., c]
        e, m = v >> mb, v & ((1 << mb) - 1)
        out[..., c] = np.where(e == 0, m / (1 << mb) * 2.0 ** -14, (1 + m / (1 << mb)) * 2.0 ** (e - 15.0))
    return out, codes


@pytest.mark.parametrize("sh_order,shadow,fill", [(1, False, 0), (2, True, 2)])
def test_indirect_specular_frame(cuda_device, sh_order, shadow, fill):
    import torch
    kw = dict(width=320, height=180, rsm_res=256, read_lod=1, cav_resolution=32, sh_order=sh_order, max_caches=16384)
    if shadow:
        kw.update(indirect_shadow=True, voxel_resolution=64, shadow_lod=1)
    wl = workloads.atrium(**kw).build()
    drv.pack_specular(wl.constant, wl.max_caches, 16)
    rm = _rough_metal(wl)
    g = workloads.DeviceFrame(wl, indirect_specular=True, specular_fill_holes_level=fill)
    g.ctx.set_constant(wl.constant)
    rm_d = torch.from_numpy(rm).cuda()
    g.ctx.bind_gbuffer_material(rm_d)
    g.prepare_inputs()
    g.frame()  # allocate -> light (+ environment maps) -> PrepareSpecularEnvmaps -> apply
    torch.cuda.synchronize()
    n = g.ctx.active_cache_count()[0]

    o = OracleFrame(wl).prepare_inputs().allocate()
    assert o.count == n and n > 500
    eo = o.alloc["entries"].copy()
    mo = orc.light_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.spot_lights, o.vpls, o.blocks, o.chain, eo, n,
                                   wl.sh_order, wl.indirect_shadow)
    e = g.ctx.read_entries(n)
    ok, ratio = close(e[:, 4:], eo[:n, 4:])
    assert ok, "SH with the flux early-out: worst |err|/tol = %.3f" % ratio
    orc.specular_mips(wl.constant, n, mo)
    orc.specular_fill_holes(wl.constant, n, fill, mo)
    mg = g.ctx.read_specular_mips()
    assert mg.shape == mo.shape
    total, S = wl.constant.SpecularEnvmapTotalSize, 16
    # level 0 before hole filling is what the light pass wrote; with fill > 0 it also holds pushed-down values —
    # compare the final chain, level by level, maps' interiors and spill borders separately
    off = 0
    for l in range(int(np.log2(S)) + 1):
        r, per = total >> l, S >> l
        a, ca = _unpack(mg[off:off + r * r].reshape(r, r))
        b, cb = _unpack(mo[off:off + r * r].reshape(r, r))
        yy, xx = np.mgrid[0:r, 0:r]
        border = (xx % per == 0) | (yy % per == 0) if per > 1 else np.ones((r, r), bool)
        # maps of the caches only: the padding invocations of the reference's last 64-cache group (ids >= count)
        # also store into "their" maps — positions read out of range — which nothing ever samples; not reproduced
        cid = (yy // per) * (total // S) + (xx // per)
        used = (b.sum(-1) > 0) & (cid < n)
        assert used.sum() > 50, l
        inner = used & ~border
        if inner.any():
            same = np.all(ca[inner] == cb[inner], -1)
            assert same.mean() >= 0.99, (l, same.mean())
            assert np.quantile(np.abs(ca[inner] - cb[inner]).max(-1), 0.9999) <= 4 + 2 * l, (l, np.abs(ca[inner] - cb[inner]).max())
        edge = used & border
        if edge.any():
            rel = np.abs(a[edge] - b[edge]).sum(-1) / np.maximum(b[edge].sum(-1), 1e-6)
            assert np.quantile(rel, 0.99) <= 0.25, (l, np.quantile(rel, 0.99))
        off += r * r
    img_o = orc.apply_caches_specular(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal,
                                      wl.diffuse, rm, o.alloc["atlas"], eo, mo)
    img = g.out32.cpu().numpy()
    assert np.array_equal(img[..., 3], img_o[..., 3])
    err = np.abs(img[..., :3].astype(np.float64) - img_o[..., :3])
    tol = 1e-4 + 2e-2 * np.maximum(np.abs(img[..., :3]), np.abs(img_o[..., :3]))
    assert np.mean(err <= tol) >= 0.999, np.mean(err <= tol)
    assert err.max() <= 20 * tol.max()
    # the specular term is there
    plain = o.apply(entries=eo)
    assert np.abs(plain[..., :3] - img_o[..., :3]).max() > 1e-4
    g.close()


def test_indirect_specular_needs_its_inputs(cuda_device):
    wl = workloads.cornell(width=64, height=64).build()
    g = workloads.DeviceFrame(wl, indirect_specular=True)
    g.prepare_inputs()
    with pytest.raises(drv.DrvError):  # Constant block without the specular fields
        g.frame()
    drv.pack_specular(wl.constant, wl.max_caches, 16)
    g.ctx.set_constant(wl.constant)
    with pytest.raises(drv.DrvError) as e:  # no roughness / metallic plane bound
        g.frame()
    assert e.value.status == abi.DRV_ERR_NOT_BOUND
    g.close()


def test_renderer_mirror_with_indirect_specular(cuda_device):
    """The reference-shaped setters (renderer.hpp:80-106) drive the same frame as the context-level calls: Draw with
    SetIndirectSpecular(true) runs PrepareSpecularEnvmaps between the light and the apply pass (renderer.cpp:557-558)."""
    import torch
    wl = workloads.atrium(width=320, height=180, rsm_res=256, read_lod=1, cav_resolution=32, sh_order=1,
                          max_caches=16384).build()
    rm = torch.from_numpy(_rough_metal(wl)).cuda()
    # context-level frame
    drv.pack_specular(wl.constant, wl.max_caches, 8)
    g = workloads.DeviceFrame(wl, indirect_specular=True, specular_per_cache_size=8, specular_fill_holes_level=1)
    g.ctx.set_constant(wl.constant)
    g.ctx.bind_gbuffer_material(rm)
    g.prepare_inputs()
    hdr_ref = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    g.ctx.draw(hdr_ref, abi.DRV_HDR_RGBA16F_ADD)
    torch.cuda.synchronize()
    # the mirror
    scene = drv.Scene(lights=wl.lights, bbox_min=wl.bbox[0], bbox_max=wl.bbox[1])
    r = drv.Renderer(scene, (wl.width, wl.height))
    r.SetCAVCascades(wl.cav_cascades, wl.cav_resolution)
    for i, s in enumerate(wl.cascade_sizes):
        r.SetCAVCascadeWorldSize(i, s)
    r.SetCAVCascadeTransitionSize(wl.transition)
    r.SetIndirectShadow(False)
    r.SetMaxCacheCount(wl.max_caches)
    assert r.GetIndirectSpecular() is False and r.GetPerCacheSpecularEnvMapSize() == 16  # renderer.cpp:43-49
    r.SetIndirectSpecular(True)
    r.SetSpecularEnvMapHoleFillLevel(9)
    assert r.GetSpecularEnvMapHoleFillLevel() == 4  # clamped to log2(16), renderer.hpp:101
    r.SetPerCacheSpecularEnvMapSize(8)
    assert r.GetSpecularEnvMapHoleFillLevel() == 3  # re-clamped, renderer.cpp:459
    r.SetSpecularEnvMapHoleFillLevel(1)
    with pytest.raises(NotImplementedError):
        r.SetSpecularEnvMapDirectWrite(False)
    dev = [torch.from_numpy(a).cuda() for a in (wl.depth, wl.normal, wl.diffuse)]
    rsm = [torch.from_numpy(a).cuda() for a in wl.rsms[0]]
    torch.cuda.synchronize()
    r.BindGBuffer(*dev, roughnessMetallic=rm)
    r.BindShadowMap(0, *rsm)
    hdr = r.Draw(wl.camera, False, 0.0)
    torch.cuda.synchronize()
    assert r.m_constant.SpecularEnvmapPerCacheSize_Texel == 8
    assert torch.equal(hdr, hdr_ref)
    assert float(hdr.float().abs().max()) > 0
    g.close()
