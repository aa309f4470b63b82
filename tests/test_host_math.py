"""The packers of include/drv_math.h against independent numpy restatements of
Renderer::Update*UBO / PrepareLights (renderer.cpp:290-431, 664-725) and the ei maths."""
import math

import numpy as np

import dynamicradiancevolume_b200 as drv


def _m(a):
    return np.array(list(a), dtype=np.float64).reshape(4, 4)


def _camera(pos, target, up=(0, 1, 0)):
    pos, target, up = (np.asarray(v, np.float64) for v in (pos, target, up))
    z = (target - pos) / np.linalg.norm(target - pos)
    x = np.cross(z, up); x /= np.linalg.norm(x)
    y = np.cross(x, z)
    look = np.eye(4); look[0, :3], look[1, :3], look[2, :3] = x, y, z
    tr = np.eye(4); tr[:3, 3] = -pos
    return look @ tr


def _perspective_dx(fovy, aspect, n, f):
    h = math.tan(math.pi * 0.5 - fovy / 2.0)
    m = np.zeros((4, 4)); m[0, 0] = h / aspect; m[1, 1] = h; m[2, 2] = f / (f - n); m[2, 3] = -n * f / (f - n); m[3, 2] = 1
    return m


def test_constant_block_values_and_signs():
    c = drv.pack_constant(1920, 1080, 128, 64, 2, 65536)
    pi = math.pi
    np.testing.assert_allclose(c.ShCosLobeFactor0, math.sqrt(pi) / 2, rtol=1e-6)
    np.testing.assert_allclose(c.ShCosLobeFactor1, math.sqrt(pi / 3), rtol=1e-6)          # positive, renderer.cpp:297
    np.testing.assert_allclose(c.ShCosLobeFactor2n2_p1_n1, -math.sqrt(15 * pi) / 8, rtol=1e-6)  # negative (sic), :298
    np.testing.assert_allclose(c.ShCosLobeFactor20, math.sqrt(5 * pi) / 16, rtol=1e-6)
    np.testing.assert_allclose(c.ShCosLobeFactor2p2, math.sqrt(15 * pi) / 16, rtol=1e-6)
    np.testing.assert_allclose(c.ShEvaFactor0, 1 / (2 * math.sqrt(pi)), rtol=1e-6)
    np.testing.assert_allclose(c.ShEvaFactor1, math.sqrt(3) / (2 * math.sqrt(pi)), rtol=1e-6)
    np.testing.assert_allclose(c.ShEvaFactor2n2_p1_n1, math.sqrt(15 / (4 * pi)), rtol=1e-6)
    np.testing.assert_allclose(c.ShEvaFactor20, math.sqrt(5 / (16 * pi)), rtol=1e-6)
    np.testing.assert_allclose(c.ShEvaFactor2p2, math.sqrt(15 / (16 * pi)), rtol=1e-6)
    assert list(c.BackbufferResolution) == [1920, 1080]
    assert (c.VoxelResolution, c.AddressVolumeResolution, c.NumAddressVolumeCascades, c.MaxNumLightCaches) == (128, 64, 2, 65536)


def test_per_frame_matrices():
    cam = drv.Camera(position=(0, 2.5, 5), direction=(0, -2.5, -5), aspect_ratio=16 / 9)
    pf = drv.pack_per_frame(cam, 1.5)
    d = np.array(cam.direction, np.float64); d /= np.linalg.norm(d)
    view = _camera(cam.position, np.array(cam.position) + d)
    proj = _perspective_dx(math.radians(60.0), 16 / 9, 1000.0, 0.1)  # (far, near) swapped, camera.hpp:28
    np.testing.assert_allclose(_m(pf.Projection), proj, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_m(pf.ViewProjection), proj @ view, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_m(pf.InverseView) @ view, np.eye(4), atol=1e-5)
    np.testing.assert_allclose(_m(pf.InverseViewProjection) @ _m(pf.ViewProjection), np.eye(4), atol=2e-3)
    np.testing.assert_allclose(list(pf.CameraDirection), d, rtol=1e-6)
    assert pf.PassedTime == 1.5
    # reversed Z: a point at the near plane has depth 1, far away -> 0
    p = np.array([*(np.array(cam.position) + d * 0.1), 1.0])
    clip = _m(pf.ViewProjection) @ p
    assert abs(clip[2] / clip[3] - 1.0) < 1e-4


def test_volume_info_snapping_and_decision_boxes():
    cam = drv.Camera(position=(0.3, 2.55, 1.97))
    vi = drv.pack_volume_info(cam, (-5.5, -0.5, -5.5), (5.5, 7.5, 7.5), 128, 64, [8.0, 16.0], 2.0)
    # voxel cube: bbox +-0.001 padded to a cube on the largest extent (renderer.cpp:351-359)
    np.testing.assert_allclose(list(vi.VolumeWorldMin), [-5.501, -0.501, -5.501], rtol=1e-6)
    ext = 13.002
    np.testing.assert_allclose(list(vi.VolumeWorldMax), [-5.501 + ext, -0.501 + ext, 7.501], rtol=1e-5)
    np.testing.assert_allclose(vi.VoxelSizeInWorld, ext / 128, rtol=1e-5)
    assert vi.CAVTransitionZoneSize == 2.0
    for i, size in enumerate((8.0, 16.0)):
        c = vi.AddressVolumeCascades[i]
        voxel = size / 64
        assert c.WorldVoxelSize == np.float32(voxel)
        snapped = np.round(np.array(cam.position) / voxel) * voxel
        np.testing.assert_allclose(list(c.Min), snapped - size / 2, atol=1e-5)
        np.testing.assert_allclose(list(c.Max), snapped + size / 2, atol=1e-5)
        np.testing.assert_allclose(list(c.DecisionMin), np.array(cam.position) - size / 2 + 1.5 * voxel, atol=1e-5)
        np.testing.assert_allclose(list(c.DecisionMax), np.array(cam.position) + size / 2 - 1.5 * voxel, atol=1e-5)


def test_spot_light_block_known_answers():
    # shadow LOD 2 and R = 128 with a 30 degree half angle => SuperValWidth = (2 sin30 / 128) * 4 = 0.03125
    l = drv.Light(intensity=(100, 100, 100), position=(0, 1.7, 3.3), direction=(0, 0, -1), halfAngle=math.radians(30),
                  rsmResolution=1024, rsmReadLod=3, indirectShadowComputationLod=2)
    s = drv.pack_spot_light(l)
    assert (s.RSMRenderResolution, s.RSMReadResolution) == (1024, 128)
    np.testing.assert_allclose(s.ValAreaFactor, (2 * math.sin(math.radians(30))) ** 2 / 128 ** 2, rtol=1e-5)
    np.testing.assert_allclose(s.IndirectShadowComputationSuperValWidth, 0.03125, rtol=1e-5)
    assert s.IndirectShadowComputationSampleInterval == 16 and s.IndirectShadowComputationBlockSize == 4.0
    np.testing.assert_allclose(s.IndirectShadowSamplingOffset, 0.5 + math.sqrt(2) * 4 / 2, rtol=1e-6)
    np.testing.assert_allclose(s.LightCosHalfAngle, math.cos(math.radians(30)), rtol=1e-6)
    view = _camera(l.position, np.array(l.position) + np.array(l.direction))
    proj = _perspective_dx(2 * l.halfAngle, 1.0, 10000.0, 0.1)
    np.testing.assert_allclose(_m(s.LightViewProjection), proj @ view, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(_m(s.InverseLightViewProjection) @ _m(s.LightViewProjection), np.eye(4), atol=5e-2)
    # non power-of-two RSM sizes round up (renderer.cpp:693-697)
    s2 = drv.pack_spot_light(drv.Light(rsmResolution=600, rsmReadLod=2))
    assert (s2.RSMRenderResolution, s2.RSMReadResolution) == (1024, 256)


def test_default_cascade_sizes():
    assert drv.default_cascade_world_sizes(4) == [4.0, 8.0, 16.0, 32.0]  # renderer.cpp:1181-1187


def test_write_pfm_is_the_reference_format(tmp_path):
    """drv_write_pfm (pure host code) == WritePfm, rendering/hdrimage.cpp:6-32: "PF\\n", "<w> <h>\\n", "-1.000000\\n",
    then the RGB floats of the RGBA image in memory order."""
    import ctypes as C
    lib = drv.load()
    w, h = 5, 3
    rgba = np.arange(w * h * 4, dtype=np.float32).reshape(h, w, 4) * 0.25 - 3.0
    path = str(tmp_path / "t.pfm")
    assert lib.drv_write_pfm(path.encode(), rgba.ctypes.data_as(C.c_void_p), w, h) == 0
    raw = open(path, "rb").read()
    head = b"PF\n5 3\n-1.000000\n"
    assert raw[:len(head)] == head
    assert np.array_equal(np.frombuffer(raw[len(head):], np.float32).reshape(h, w, 3), rgba[..., :3])
    assert lib.drv_write_pfm(b"/nonexistent-dir/x.pfm", rgba.ctypes.data_as(C.c_void_p), w, h) != 0
