"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/drv_gi.h declares."""
import ctypes as C
import os
import re

import pytest

import dynamicradiancevolume_b200 as drv
from dynamicradiancevolume_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "drv_gi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = drv.load()
    names = _declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libdrv_gi.so does not export %s" % n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, "binding and header drifted: %s" % (set(names) ^ bound)


def test_host_library_exports_the_packers():
    """libdrv_host.so = host_pack.cpp compiled without CUDA: same drv_pack_* symbols, same results as libdrv_gi.so."""
    host, full = _lib.load_host(), drv.load()
    assert [s[0] for s in _lib.HOST_SYMBOLS] == ["drv_pack_constant", "drv_pack_specular", "drv_pack_per_frame",
                                                 "drv_pack_volume_info", "drv_pack_spot_light"]
    a, b = abi.Constant(), abi.Constant()
    host.drv_pack_constant(C.byref(a), 1920, 1080, 128, 64, 2, 65536)
    full.drv_pack_constant(C.byref(b), 1920, 1080, 128, 64, 2, 65536)
    assert bytes(a) == bytes(b)


def test_struct_sizes_match_std140():
    assert C.sizeof(abi.Constant) == 80
    assert C.sizeof(abi.PerFrame) == 288
    assert C.sizeof(abi.VolumeInfo) == 288
    assert C.sizeof(abi.SpotLight) == 224
    assert abi.VolumeInfo.AddressVolumeCascades.offset == 32
    assert abi.SpotLight.LightViewProjection.offset == 64
    assert abi.SpotLight.IndirectShadowSamplingOffset.offset == 220
    assert abi.Constant.BackbufferResolution.offset == 40 and abi.Constant.MaxNumLightCaches.offset == 60


def test_version_and_stage_names():
    lib = drv.load()
    assert b"sm_100a" in lib.drv_version()
    assert [lib.drv_stage_name(i).decode() for i in range(len(abi.STAGE_NAMES))] == abi.STAGE_NAMES
    assert lib.drv_microbench_count() >= 5


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(drv.DrvError) as e:
        drv.Context(width=64, height=64)
    assert e.value.status == abi.DRV_ERR_NO_DEVICE


def test_invalid_config_rejected():
    lib = drv.load()
    cfg = abi.Config()
    h = C.c_void_p()
    assert lib.drv_create(C.byref(cfg), C.byref(h)) == abi.DRV_ERR_INVALID
    assert lib.drv_create(None, C.byref(h)) == abi.DRV_ERR_INVALID


def test_shard_range_partitions_on_64_entry_boundaries():
    for count in (0, 1, 63, 64, 65, 1000, 6210, 65536):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = drv.shard_range(count, r, world)
                assert b == prev and e >= b
                assert b % 64 == 0 or b == count
                prev = e
            assert prev == count
