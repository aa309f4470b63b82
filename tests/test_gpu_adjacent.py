"""GPU parity of the rows SURVEY.md 8(f) ranks next to the hot path — RSM fill (f1), voxel cone-traced ambient
occlusion (f2), tonemap + PFM (f3) — through the C-ABI against the CPU oracle."""
import math
import os

import numpy as np
import pytest

import dynamicradiancevolume_b200 as drv
import workloads
from dynamicradiancevolume_b200 import abi
from oracle import binding as orc
from oracle.frame import OracleFrame, close

pytestmark = pytest.mark.gpu


def _halfs(a):
    return np.ascontiguousarray(a).view(np.float16).astype(np.float32)


@pytest.mark.parametrize("res,ragged", [(64, False), (256, True)])
def test_fill_rsm_matches_oracle(cuda_device, res, ragged):
    """fillrsm.frag:32-61 on synthetic fragment attributes: flux and depthLinSq halfs bit-exact, packed normals
    within one int16 code; uncovered texels keep the clear value; the result is bound and feeds VPL generation."""
    import torch
    wl = workloads.cornell(rsm_res=res, read_lod=0).build()
    light = wl.spot_lights[0]
    rng = np.random.default_rng(5)
    lp = np.array(light.LightPosition[:3], np.float32)
    ld = np.array(light.LightDirection[:3], np.float32)
    # points scattered in front of the light, inside and outside its cone
    dirs = rng.normal(size=(res, res, 3)).astype(np.float32) * 0.6 + ld
    pos = (lp + dirs * rng.uniform(0.5, 6.0, size=(res, res, 1)).astype(np.float32)).astype(np.float32)
    nrm = (rng.normal(size=(res, res, 3)) * rng.uniform(0.2, 3.0, size=(res, res, 1))).astype(np.float32)
    nrm[0, 0] = (0.0, 2.0, 0.0)   # x == 0 branch of PackNormal16I
    nrm[0, 1] = (0.0, -1.0, 0.0)
    base = rng.uniform(0.0, 1.0, size=(res, res, 3)).astype(np.float32)
    cov = (rng.uniform(size=(res, res)) > 0.25).astype(np.uint8) if ragged else None
    fo, no, do = orc.fill_rsm(light, pos, nrm, base, cov)
    ctx = drv.Context(**wl.context_kwargs())
    ctx.set_constant(wl.constant); ctx.set_per_frame(wl.per_frame); ctx.set_volume_info(wl.volume)
    ctx.set_light_count(1); ctx.set_spot_light(0, light)
    t = [torch.from_numpy(a).cuda() for a in (pos, nrm, base)]
    tc = None if cov is None else torch.from_numpy(cov).cuda()
    torch.cuda.synchronize()
    ctx.fill_rsm(0, t[0], t[1], t[2], tc)
    torch.cuda.synchronize()
    b = ctx.buffers()
    # the RSM the call produced and bound
    flux = ctx._read(b.rsm_flux0[0], res * res * 8).view(np.uint16).reshape(res, res, 4)
    nor = ctx._read(b.rsm_normal0[0], res * res * 4).view(np.int16).reshape(res, res, 2)
    dep = ctx._read(b.rsm_depth0[0], res * res * 4).view(np.uint16).reshape(res, res, 2)
    assert np.array_equal(flux[..., :3], fo[..., :3])
    assert np.array_equal(dep, do)
    dn = np.abs(nor.astype(np.int32) - no.astype(np.int32))
    dn[..., 0] = np.minimum(dn[..., 0], 65536 - dn[..., 0])  # the azimuth code wraps at +-pi
    assert dn.max() <= 1, dn.max()
    assert fo[..., :3].max() > 0
    if cov is not None:
        assert not flux[cov == 0].any() and not dep[cov == 0].any() and not nor[cov == 0].any()
    # and the bound RSM drives the rest of the path
    ctx.light_caches()
    vp = ctx.read_vpls(0, res * res)
    vo = orc.generate_vpls(light, fo, no, do)
    assert np.array_equal(vp["Flux"], vo["Flux"])
    ctx.close()


def test_cone_trace_ao_matches_oracle(cuda_device):
    """ambientocclusion.frag:25-89 on the atrium at 480x270 with a 64^3 voxel volume."""
    import torch
    wl = workloads.atrium(width=480, height=270, rsm_res=64, read_lod=0, sh_order=1, indirect_shadow=True,
                          voxel_resolution=64).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    out = torch.full((wl.height, wl.width), -7.0, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    g.ctx.cone_trace_ao(out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    o = OracleFrame(wl).prepare_inputs()
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain)
    ref = orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal,
                            out=np.full((wl.height, wl.width), -7.0, np.float32))
    assert np.array_equal(got == -7.0, ref == -7.0)          # the same pixels are discarded
    covered = ref != -7.0
    assert covered.any() and (ref[covered] < 0.999).any() and ref[covered].min() >= 0.0
    err = np.abs(got - ref)[covered]
    # continuous maths on a march whose stop test (coneWeight < 0.99) can flip on a 1e-7 difference: a flipped
    # step moves the result by at most 0.01 * (pi / 4) / 6 = 1.3e-3
    assert err.max() <= 2e-3, err.max()
    assert (err > 1e-4).mean() < 1e-3
    g.close()


def test_cone_trace_ao_needs_the_record_chain(cuda_device):
    import torch
    wl = workloads.cornell().build()
    g = workloads.DeviceFrame(wl)
    out = torch.zeros(wl.height, wl.width, dtype=torch.float32, device="cuda")
    with pytest.raises(drv.DrvError) as e:
        g.ctx.cone_trace_ao(out)
    assert e.value.status == abi.DRV_ERR_NOT_BOUND
    g.close()


def test_tonemap_and_pfm(cuda_device, tmp_path):
    """tonemapping.frag:21-31 on the frame's RGBA16F target, and SaveToPFM / WritePfm (hdrimage.cpp:6-32)."""
    import torch
    wl = workloads.config(0).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    g.frame(hdr, abi.DRV_HDR_RGBA16F_ADD)
    ldr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float32, device="cuda")
    exposure, l_max = 2.5, 1.2
    g.ctx.tonemap(hdr, ldr, exposure, l_max)
    torch.cuda.synchronize()
    h = hdr.float().cpu().numpy()
    ref = orc.tonemap(h, exposure, np.float32(math.log2(l_max + 1.0)))
    got = ldr.cpu().numpy()
    ok, ratio = close(got[..., :3], ref)
    assert ok, ratio
    assert np.all(got[..., 3] == 1.0) and ref.max() > 0.01
    path = str(tmp_path / "frame.pfm")
    g.ctx.save_to_pfm(hdr, path)
    raw = open(path, "rb").read()
    header = b"PF\n%d %d\n-1.000000\n" % (wl.width, wl.height)
    assert raw.startswith(header)
    body = np.frombuffer(raw[len(header):], np.float32).reshape(wl.height, wl.width, 3)
    assert np.array_equal(body, h[..., :3])
    g.close()
