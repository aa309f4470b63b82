"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle and the scene generator still reproduce them bit for bit (regression pin).
GPU: libdrv_gi through the C-ABI against the stored vectors — allocation / voxel set bit-exact, SH and radiance
within 1e-3 relative / 1e-5 absolute."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)

CASES = sorted(mg.cases().keys())


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    gold = _load(name)
    wl = mg.cases()[name]
    out = mg.run(wl)
    assert str(out["digest"]) == str(gold["digest"]), "the synthetic inputs drifted (scenes/ or the packers changed)"
    for k in gold.files:
        if gold[k].dtype.kind in "US":
            assert str(out[k]) == str(gold[k]), k
        else:
            assert np.array_equal(out[k], gold[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_golden(cuda_device, name):
    import torch
    import workloads
    from oracle.frame import close
    gold = _load(name)
    wl = mg.cases()[name].build()
    assert mg.input_digest(wl) == str(gold["digest"])
    g = workloads.DeviceFrame(wl)
    g.prepare_inputs()
    g.frame()
    torch.cuda.synchronize()
    n = g.ctx.active_cache_count()[0]
    assert n == len(gold["cell_ids"])
    atlas = g.ctx.read_atlas()
    R = wl.cav_resolution
    zz, yy, xx = np.nonzero(atlas)
    ids = np.sort((xx % R) + yy * R + zz * R * R + (xx // R) * R ** 3).astype(np.int32)
    assert np.array_equal(ids, gold["cell_ids"])
    e = g.ctx.read_entries(n)
    assert np.array_equal(e[:, :4], gold["entries"][:, :4])
    ok, ratio = close(e[:, 4:], gold["entries"][:, 4:])
    assert ok, ratio
    img = g.out32.cpu().numpy()
    assert np.array_equal(np.packbits(img[..., 3] > 0), gold["shaded"])
    ok, ratio = close(img[..., :3], gold["image"])
    assert ok, ratio
    if wl.indirect_shadow:
        assert hashlib.sha256(g.ctx.read_voxel_chain().tobytes()).hexdigest() == str(gold["voxel_chain_sha256"])
        assert np.array_equal(np.packbits(g.ctx.read_voxel_target() > 0), gold["voxel_set"])
        nb = len(gold["shadow_blocks"])
        assert np.array_equal(g.ctx.read_shadow_blocks(0, nb).view(np.float32).reshape(-1, 4), gold["shadow_blocks"])
    v = g.ctx.read_vpls(0, 64).view(np.float32).reshape(-1, 12)
    assert np.array_equal(v[:, :4], gold["vpl_head"][:, :4])
    # rows next to the path
    if wl.indirect_shadow:
        ao = torch.zeros(wl.height, wl.width, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        g.ctx.cone_trace_ao(ao)
        torch.cuda.synchronize()
        assert np.abs(ao.cpu().numpy() - gold["ao"]).max() <= 2e-3
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    ldr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    g.ctx.apply_caches(hdr, 0)
    g.ctx.tonemap(hdr, ldr, 2.0, 1.2)
    torch.cuda.synchronize()
    ok, ratio = close(ldr.cpu().numpy()[..., :3], gold["tonemap"], rtol=2e-3, atol=1e-4)  # through the RGBA16F target
    assert ok, ratio
    g.close()
