"""Whole-frame GPU parity on the BASELINE.json configurations: allocate -> light -> apply through the
C-ABI (``drv_draw``) against the CPU oracle, plus the host-buffer entry point and the Renderer mirror."""
import numpy as np
import pytest

import dynamicradiancevolume_b200 as drv
import workloads
from dynamicradiancevolume_b200 import abi
from oracle.frame import OracleFrame, close

pytestmark = pytest.mark.gpu


def _run(wl, **kw):
    import torch
    wl.build()
    g = workloads.DeviceFrame(wl, **kw)
    o = OracleFrame(wl).prepare_inputs()
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    img_o = o.frame()
    n, overflow, oob = g.ctx.active_cache_count()
    assert (n, overflow) == (o.count, 0)
    e = g.ctx.read_entries(n)
    img = g.out32.cpu().numpy()
    return g, o, e, img, img_o


def _check(g, o, e, img, img_o):
    n = o.count
    assert np.array_equal(e[:, :4], o.entries[:n, :4])
    ok, ratio = close(e[:, 4:], o.entries[:n, 4:])
    assert ok, "SH: worst |err|/tol = %.3f" % ratio
    assert np.array_equal(img[..., 3], img_o[..., 3])
    ok, ratio = close(img[..., :3], img_o[..., :3])
    assert ok, "radiance: worst |err|/tol = %.3f" % ratio
    assert img_o[..., :3].max() > 1e-3


def test_config0_cornell_sh1_unshadowed(cuda_device):
    """BASELINE configs[0]: Cornell 512x512, 64x64 RSM (4k VPLs), 1 x 32^3, SH1, no indirect shadow."""
    g, o, e, img, img_o = _run(workloads.config(0))
    _check(g, o, e, img, img_o)
    g.close()


def test_config0_from_rsm_mip(cuda_device):
    """Same, but the RSM is rendered at 1024^2 and read at LOD 4 (the reference's defaults, scene/light.hpp:12):
    VPL normals now come from the GPU's own mip chain (1-LSB differences in the int16 normal code are allowed)."""
    g, o, e, img, img_o = _run(workloads.cornell(rsm_res=1024, read_lod=4))
    _check(g, o, e, img, img_o)
    g.close()


def test_config1_atrium_1080p(cuda_device):
    """BASELINE configs[1]: 1920x1080, 128^2 RSM read (16k VPLs), 2 x 64^3 with transitions, SH1, unshadowed."""
    g, o, e, img, img_o = _run(workloads.config(1))
    _check(g, o, e, img, img_o)
    g.close()


def test_config2_atrium_shadow_sh2(cuda_device):
    """BASELINE configs[2]: C2 + 128^3 voxels + mip chain, cone-traced visibility, SH2."""
    g, o, e, img, img_o = _run(workloads.config(2))
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain)
    _check(g, o, e, img, img_o)
    g.close()


def test_config3_reduced_4_lights_4_cascades(cuda_device):
    """BASELINE configs[3] at reduced resolution (960x540, 4 x 64^3, 4 lights x 64^2 read): SH2 + shadows."""
    wl = workloads.config(3, width=960, height=540, cav_resolution=64, rsm_res=256, read_lod=2, max_caches=1 << 17,
                          voxel_resolution=64)
    g, o, e, img, img_o = _run(wl)
    _check(g, o, e, img, img_o)
    g.close()


def test_reference_default_settings(cuda_device):
    """The settings the reference's author ran by default: 1080p, 4096 VPLs, 3 cascades x 32^3 with
    transitions, 128^3 voxels, SH1 + cone-traced shadows at LOD 2, 16384-cache capacity."""
    g, o, e, img, img_o = _run(workloads.config(5))
    _check(g, o, e, img, img_o)
    g.close()


def test_draw_to_host_matches_device_path(cuda_device):
    """drv_upload_* + drv_draw_to_host (the end-to-end entry point bench.py times) against the device path."""
    import torch
    wl = workloads.config(0).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    ref = g.out32.cpu().numpy()
    ctx = drv.Context(**wl.context_kwargs())
    ctx.set_constant(wl.constant); ctx.set_per_frame(wl.per_frame); ctx.set_volume_info(wl.volume)
    ctx.set_light_count(1); ctx.set_spot_light(0, wl.spot_lights[0])
    pin = lambda a: torch.from_numpy(a).pin_memory()
    d, n, c = pin(wl.depth), pin(wl.normal), pin(wl.diffuse)
    f, rn, rd = (pin(a) for a in wl.rsms[0])
    ctx.upload_gbuffer(d, n, c)
    ctx.upload_rsm(0, f, rn, rd)
    ctx.prepare_rsm(0)
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16).pin_memory()
    ctx.draw_to_host(hdr)
    out = hdr.float().numpy()
    expect = ref[..., :3].astype(np.float16).astype(np.float32)
    assert np.array_equal(out[..., :3], expect)
    assert ctx.kernel_launches() > 0
    ctx.close()
    g.close()


@pytest.mark.parametrize("cfg", [0, 2])
def test_draw_host_frame_pipelined(cuda_device, cfg):
    """drv_draw_host_frame (banded H2D / apply / D2H overlap) gives the same image as the unpipelined host path."""
    import torch
    wl = workloads.config(cfg).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    g.frame(hdr, abi.DRV_HDR_RGBA16F_ADD)
    torch.cuda.synchronize()
    ref = hdr.cpu()
    pin = lambda a: torch.from_numpy(a).pin_memory()
    gb = [pin(a) for a in (wl.depth, wl.normal, wl.diffuse)]
    rsms = [[pin(a) for a in r] for r in wl.rsms]
    out = torch.full((wl.height, wl.width, 4), 7.0, dtype=torch.float16).pin_memory()
    for bands in (0, 1, 5, 32):
        out.fill_(7.0)
        g.ctx.draw_host_frame(gb[0], gb[1], gb[2], rsms, out, bands)
        assert torch.equal(out, ref), bands
    g.close()


@pytest.mark.parametrize("cfg", [0, 2])
def test_draw_frame_matches_serial_order(cuda_device, cfg):
    """drv_draw_frame (light side || camera side, fused clear, CUDA-graph replay) against the reference's serial
    order prepare_rsm -> clear -> drv_draw: bit-identical images and entries, eager and replayed."""
    import torch
    wl = workloads.config(cfg).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    g.frame(hdr, abi.DRV_HDR_RGBA16F_ADD)
    torch.cuda.synchronize()
    ref = hdr.cpu()
    n = g.ctx.active_cache_count()[0]
    ref_entries = g.ctx.read_entries(n)
    out = torch.full((wl.height, wl.width, 4), 7.0, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    g.ctx.bind_scene(g.tris, None, 1.0)
    vox = abi.DRV_FRAME_VOXELIZE if wl.indirect_shadow else 0
    for flags in (abi.DRV_FRAME_PREPARE_RSM, abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH,
                  abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH | vox):
        for rep in range(4):  # graph mode: eager, record + replay, replay, replay
            out.fill_(7.0)
            torch.cuda.synchronize()
            launches = g.ctx.kernel_launches()
            g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, flags)
            torch.cuda.synchronize()
            assert g.ctx.kernel_launches() > launches
            assert torch.equal(out.cpu(), ref), (flags, rep)
            assert np.array_equal(g.ctx.read_entries(n), ref_entries), (flags, rep)
    # a changed uniform block invalidates the recorded graph: the next frame must see the new camera
    import copy
    pf = copy.copy(wl.per_frame)
    g.ctx.set_per_frame(pf)
    g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)
    g.close()


def test_draw_frame_edge_cases(cuda_device):
    """drv_draw_frame on degenerate frames, eager and replayed: an empty G-buffer (no caches: every kernel must cope with
    a zero count read on the device), a frame after it with geometry again (stale scan / queue state must not leak),
    capacity overflow, and an RSM whose flux is zero everywhere (empty live-VPL list)."""
    import torch
    wl = workloads.cornell(width=160, height=96, rsm_res=64, read_lod=1, sh_order=2, indirect_shadow=True,
                           voxel_resolution=32).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    g.ctx.bind_scene(g.tris, None, 1.0)
    flags = abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_VOXELIZE | abi.DRV_FRAME_GRAPH
    out = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    ref = torch.zeros_like(out)
    torch.cuda.synchronize()
    g.frame(ref, abi.DRV_HDR_RGBA16F_ADD)
    torch.cuda.synchronize()
    n_ref = g.ctx.active_cache_count()[0]
    depth = g.depth.clone()
    # 1. nothing visible
    g.depth.zero_()
    for _ in range(3):
        out.fill_(3.0)
        g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, flags)
        torch.cuda.synchronize()
        assert g.ctx.active_cache_count()[:2] == (0, 0)
        assert float(out.float().abs().max()) == 0.0
    # 2. geometry is back: same result as before the empty frames, from the same recorded graph
    g.depth.copy_(depth)
    for _ in range(2):
        g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, flags)
        torch.cuda.synchronize()
        assert g.ctx.active_cache_count()[0] == n_ref
        assert torch.equal(out, ref)
    # 3. an RSM without flux: no live VPLs, caches are allocated but stay dark
    flux = g.rsms[0][0].clone()
    g.rsms[0][0].zero_()
    for _ in range(2):
        g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, flags)
        torch.cuda.synchronize()
        assert g.ctx.live_vpl_counts()[0] == 0 and g.ctx.active_cache_count()[0] == n_ref
        assert float(out.float().abs().max()) == 0.0
    g.rsms[0][0].copy_(flux)
    g.ctx.draw_frame(out, abi.DRV_HDR_RGBA16F_WRITE, flags)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    g.close()
    # 4. capacity overflow through the frame call
    wl2 = workloads.cornell(max_caches=100).build()
    g2 = workloads.DeviceFrame(wl2)
    o = OracleFrame(wl2).allocate()
    out2 = torch.zeros(wl2.height, wl2.width, 4, dtype=torch.float16, device="cuda")
    for _ in range(3):
        g2.ctx.draw_frame(out2, abi.DRV_HDR_RGBA16F_WRITE, abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH)
    torch.cuda.synchronize()
    n, overflow, _ = g2.ctx.active_cache_count()
    assert n == 100 and overflow == o.alloc["overflow"] and overflow > 0
    assert np.array_equal(g2.ctx.read_atlas(), o.alloc["atlas"])
    g2.close()


def test_renderer_mirror_draw(cuda_device):
    """The reference-shaped host interface (Renderer::Draw, renderer.cpp:501-594) drives the same frame."""
    import torch
    wl = workloads.config(0, indirect_shadow=True, sh_order=2).build()
    scene = drv.Scene(lights=wl.lights, bbox_min=wl.bbox[0], bbox_max=wl.bbox[1])
    tris = torch.from_numpy(wl.triangles.reshape(-1).copy()).cuda()
    scene.entities = [(tris, None)]
    r = drv.Renderer(scene, (wl.width, wl.height))
    r.SetCAVCascades(wl.cav_cascades, wl.cav_resolution)
    r.SetCAVCascadeWorldSize(0, wl.cascade_sizes[0])
    r.SetCAVCascadeTransitionSize(wl.transition)
    r.SetIndirectDiffuseMode(drv.IndirectDiffuseMode.SH2)
    r.SetIndirectShadow(True)
    r.SetVoxelVolumeResultion(wl.voxel_resolution)
    r.SetVoxelVolumeAdaptionRate(1.0)
    r.SetMaxCacheCount(wl.max_caches)
    dev = [torch.from_numpy(a).cuda() for a in (wl.depth, wl.normal, wl.diffuse)]
    rsm = [torch.from_numpy(a).cuda() for a in wl.rsms[0]]
    torch.cuda.synchronize()
    r.BindGBuffer(*dev)
    r.BindShadowMap(0, *rsm)
    r.SetReadLightCacheCount(True)
    hdr = r.Draw(wl.camera, False, 1.0)  # dt * rate * 255 = 255 -> adaption 1: converged volume
    torch.cuda.synchronize()
    o = OracleFrame(wl).prepare_inputs()
    img_o = o.frame()
    out = hdr.float().cpu().numpy()
    ok, ratio = close(out[..., :3], img_o[..., :3], rtol=2e-3, atol=1e-4)  # RGBA16F target
    assert ok, ratio
    r.Draw(wl.camera, False, 0.0)
    assert r.GetLightCacheActiveCount() == o.count  # one frame late, renderer.cpp:960-966
    # the passes around the path through the same mirror: AO output mode, tonemap, screenshot
    from oracle import binding as orc
    ao_t = r.ConeTraceAO()
    torch.cuda.synchronize()  # the context works on its own stream
    ao = ao_t.cpu().numpy()
    ao_ref = orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal)
    assert np.abs(ao - ao_ref).max() <= 2e-3
    r.SetExposure(3.0)
    ldr_t = r.Tonemap()
    torch.cuda.synchronize()
    ldr = ldr_t.cpu().numpy()
    hdr_now = r._hdr.float().cpu().numpy()
    ok, ratio = close(ldr[..., :3], orc.tonemap(hdr_now, 3.0, np.float32(np.log2(r.GetTonemapLMax() + 1.0))))
    assert ok, ratio


def test_config3_full_size_stated_subsample(cuda_device):
    """BASELINE configs[3] AT FULL SIZE (3840x2160, 4 lights x 128^2 = 65 536 VPLs, 4 cascades x 128^3, SH2 +
    cone-traced shadows through the 128^3 chain): allocation bit-exact on the whole frame, the SH of every 64th
    entry against the oracle's gather, the whole image against the oracle's apply pass (oracle/subsample.py) —
    for the serial stage order into RGBA32F and for drv_draw_frame (graph replay) into RGBA16F."""
    import torch
    from oracle.subsample import check_frame
    wl = workloads.config(3).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    n, overflow, _ = g.ctx.active_cache_count()
    assert overflow == 0 and n > 10000
    o = OracleFrame(wl).prepare_inputs().allocate()
    entries, atlas = g.ctx.read_entries(n), g.ctx.read_atlas()
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain)
    r = check_frame(wl, entries, atlas, n, g.out32.cpu().numpy(), step=64, oracle=o)
    assert r["alloc_exact"], r
    assert r["checked_entries"] >= n // 64 and r["sh_max_abs"] > 0 and r["image_max"] > 0
    assert r["sh_ok"] and r["image_ok"], r
    # the overlapped / graph-replayed frame into the reference's RGBA16F target
    out16 = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    g.ctx.bind_scene(g.tris, None, 1.0)
    for _ in range(3):
        g.ctx.draw_frame(out16, abi.DRV_HDR_RGBA16F_WRITE,
                         abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_VOXELIZE | abi.DRV_FRAME_GRAPH)
    torch.cuda.synchronize()
    r16 = check_frame(wl, g.ctx.read_entries(n), g.ctx.read_atlas(), n, out16.float().cpu().numpy(), step=64,
                      image_is_half=True, oracle=o)
    assert r16["ok"], r16
    g.close()


def test_frame_graph_survives_a_moving_camera(cuda_device):
    """drv_set_per_frame / drv_set_volume_info every frame (an animated camera): the recorded frame graph is patched
    in place (cudaGraphExecUpdate), never re-instantiated, and every frame equals the eager frame of the same
    uniforms bit for bit — and the oracle's frame within the gate."""
    import copy
    import torch
    wl = workloads.atrium(width=640, height=360, rsm_res=256, read_lod=1, cav_resolution=32).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    out_g = torch.zeros(wl.height, wl.width, 4, dtype=torch.float32, device="cuda")
    out_e = torch.zeros_like(out_g)
    flags = abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH
    for _ in range(3):
        g.ctx.draw_frame(out_g, abi.DRV_HDR_RGBA32F_WRITE, flags)
    inst0, upd0 = g.ctx.graph_stats()
    assert inst0 == 1
    cam = copy.copy(wl.camera)
    counts = set()
    for step in range(6):
        cam.position = (0.15 * step, 2.5 + 0.05 * step, 2.0 - 0.2 * step)
        pf = drv.pack_per_frame(cam, 0.1 * step)
        vi = drv.pack_volume_info(cam, wl.bbox[0], wl.bbox[1], wl.voxel_resolution, wl.cav_resolution, wl.cascade_sizes,
                                  wl.transition)
        # the G-buffer is that of the original camera: unprojection and cascades move with the uniforms, which is
        # all this test needs (a different set of cells is allocated every frame)
        g.ctx.set_per_frame(pf)
        g.ctx.set_volume_info(vi)
        g.ctx.draw_frame(out_g, abi.DRV_HDR_RGBA32F_WRITE, flags)
        torch.cuda.synchronize()
        n = g.ctx.active_cache_count()[0]
        counts.add(n)
        e_g = g.ctx.read_entries(n)
        g.ctx.draw_frame(out_e, abi.DRV_HDR_RGBA32F_WRITE, abi.DRV_FRAME_PREPARE_RSM)  # eager, same uniforms
        torch.cuda.synchronize()
        assert g.ctx.active_cache_count()[0] == n
        assert np.array_equal(g.ctx.read_entries(n), e_g)
        assert torch.equal(out_g, out_e)
        if step == 5:
            wl2 = copy.copy(wl)
            wl2.per_frame, wl2.volume = pf, vi
            o = OracleFrame(wl2).prepare_inputs()
            img = o.frame()
            assert o.count == n
            ok, ratio = close(out_g.cpu().numpy()[..., :3], img[..., :3])
            assert ok, ratio
    inst1, upd1 = g.ctx.graph_stats()
    assert inst1 == inst0, "a uniform change re-instantiated the frame graph"
    assert upd1 - upd0 == 6
    assert len(counts) > 1
    g.close()


@pytest.mark.parametrize("res", [256, 512])
def test_config2_with_larger_voxel_volumes(cuda_device, res):
    """BASELINE configs[2] with the voxel volume at 256^3 and 512^3 (the reference's own sweep and its UI go up to
    512^3, application.cpp:345, tweakbarsetup.cpp:185): chain bit-exact, gather-ready records (175 MB / 1.3 GB, past
    the 126 MB L2) exercised by the cone pass: SH on every 16th entry within the gate, the whole image through the
    oracle's apply pass."""
    import torch
    from oracle.subsample import check_frame
    wl = workloads.config(2, voxel_resolution=res).build()
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    n = g.ctx.active_cache_count()[0]
    o = OracleFrame(wl).prepare_inputs().allocate()
    assert np.array_equal(g.ctx.read_voxel_chain(), o.chain)
    r = check_frame(wl, g.ctx.read_entries(n), g.ctx.read_atlas(), n, g.out32.cpu().numpy(), step=16, oracle=o)
    assert r["ok"] and r["sh_max_abs"] > 0, r
    g.close()
