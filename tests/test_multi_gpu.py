"""Sharded gather across GPUs (needs >= 2 CUDA devices; skipped on the 1-GPU box): fused P2P all-gather from the
gather epilogue and the NCCL collective exchange, both against the CPU oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import workloads
    from dynamicradiancevolume_b200 import abi, sharding
    wl = workloads.atrium(width=640, height=360, rsm_res=256, read_lod=2, sh_order=2, indirect_shadow=True,
                          voxel_resolution=64, num_lights=2).build()
    stream = torch.cuda.Stream(device=rank)
    g = workloads.DeviceFrame(wl, device=rank, stream=stream)
    word = torch.zeros(1, dtype=torch.int32, device="cuda:%d" % rank)
    if mode.startswith("fused"):
        sharding.connect_peers(g.ctx, rank, world)
        if mode.endswith("interleaved"):
            g.ctx.set_shard_interleave(True)  # 64-entry groups dealt round-robin instead of contiguous ranges
    else:
        g.ctx.set_shard(rank, world)
    bctx = g.ctx if mode in ("fused", "fused_interleaved") else None  # "fused": flags in peer memory; otherwise an NCCL all-reduce
    if mode.startswith("fused_frame"):
        # drv_draw_frame on every rank: replicated allocation || light side, peer barriers, sharded gather with the
        # fused all-gather, this rank's band of the apply pass; frames 1.. replay the recorded CUDA graph
        flags = abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH | abi.DRV_FRAME_APPLY_OWN_ROWS
        for it in range(4):
            with torch.cuda.stream(stream):
                g.ctx.voxelize(g.tris, None, 1.0)
                g.out32.zero_()
                g.ctx.draw_frame(g.out32, abi.DRV_HDR_RGBA32F_WRITE, flags)
            torch.cuda.synchronize()
        bands = [torch.zeros_like(g.out32) for _ in range(world)]
        dist.all_gather(bands, g.out32)
        g.out32.copy_(sum(bands))  # bands are disjoint, the rest of every image is zero
        # the same frame with the fused image gather: every rank's band lands in rank 0's RGBA16F target over NVLink
        sharding.connect_image_gather(g.ctx, rank, world)
        gflags = abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH | abi.DRV_FRAME_GATHER_IMAGE
        for it in range(3):
            with torch.cuda.stream(stream):
                g.ctx.draw_frame(None, abi.DRV_HDR_RGBA16F_WRITE, gflags)
            torch.cuda.synchronize()
        if rank == 0:
            got = g.ctx.hdr16_tensor().float()
            want = g.out32[..., :3].half().float()
            assert torch.equal(got[..., :3], want), "fused image gather differs from the banded apply"
    for it in range(0 if mode.startswith("fused_frame") else 2):  # twice: the second frame checks the cross-frame ordering of clears and peer stores
        with torch.cuda.stream(stream):
            g.prepare_inputs()
            g.ctx.allocate_caches()
            sharding.barrier(word, ctx=bctx)
            g.ctx.light_caches()
            if mode.startswith("fused"):
                sharding.barrier(word, ctx=bctx)
            else:
                n = g.ctx.active_cache_count()[0]
                sharding.exchange_entries(g.ctx.entries_tensor(), n, world)
            g.ctx.apply_caches(g.out32, abi.DRV_HDR_RGBA32F_WRITE)
        torch.cuda.synchronize()
    n = g.ctx.active_cache_count()[0]
    np.save(os.path.join(out_dir, "entries_%d.npy" % rank), g.ctx.read_entries(n))
    np.save(os.path.join(out_dir, "image_%d.npy" % rank), g.out32.cpu().numpy())
    dist.barrier()
    g.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fused", "fused_nccl_barrier", "nccl", "fused_frame", "fused_frame_interleaved",
                                  "fused_interleaved"])
def test_sharded_gather_matches_oracle(tmp_path, mode):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import workloads
    from oracle.frame import OracleFrame, close
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    wl = workloads.atrium(width=640, height=360, rsm_res=256, read_lod=2, sh_order=2, indirect_shadow=True,
                          voxel_resolution=64, num_lights=2).build()
    o = OracleFrame(wl).prepare_inputs()
    img = o.frame()
    first = np.load(tmp_path / "entries_0.npy")
    for r in range(world):
        e = np.load(tmp_path / ("entries_%d.npy" % r))
        assert e.shape[0] == o.count
        assert np.array_equal(e, first)  # every rank ends with identical bytes
        ok, ratio = close(e[:, 4:], o.entries[:o.count, 4:])
        assert ok, (r, ratio)
        im = np.load(tmp_path / ("image_%d.npy" % r))
        ok, ratio = close(im[..., :3], img[..., :3])
        assert ok, (r, ratio)
