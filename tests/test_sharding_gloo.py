"""Host-side multi-rank logic on CPU: world_size 2 and 3 over gloo. The CPU oracle stands in for the gather kernel
(test infrastructure), so this covers the shard partition and the exchange, not the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import workloads
    from dynamicradiancevolume_b200 import sharding
    from oracle.frame import OracleFrame
    wl = workloads.cornell(width=128, height=128, rsm_res=32, sh_order=2, indirect_shadow=True, voxel_resolution=32,
                           shadow_lod=1).build()
    o = OracleFrame(wl, threads=1).prepare_inputs().allocate()  # replicated, deterministic
    b, e = sharding.shard_ranges(o.count, world)[rank]
    o.light(first=b, count=e - b)  # this rank's shard only
    entries = torch.from_numpy(o.entries)
    # rows outside the own shard are still zero here
    other = np.ones(o.count, bool); other[b:e] = False
    assert not o.entries[:o.count][other][:, 4:].any()
    sharding.exchange_entries(entries, o.count, world)
    img = o.apply()
    np.save(os.path.join(out_dir, "entries_%d.npy" % rank), o.entries[:o.count])
    np.save(os.path.join(out_dir, "image_%d.npy" % rank), img)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_light_and_exchange_equals_single_rank(tmp_path, world):
    import workloads
    from oracle.frame import OracleFrame
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    wl = workloads.cornell(width=128, height=128, rsm_res=32, sh_order=2, indirect_shadow=True, voxel_resolution=32,
                           shadow_lod=1).build()
    o = OracleFrame(wl, threads=1).prepare_inputs()
    img = o.frame()
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("entries_%d.npy" % r)), o.entries[:o.count])
        assert np.array_equal(np.load(tmp_path / ("image_%d.npy" % r)), img)


def test_shard_ranges_cover_and_balance():
    from dynamicradiancevolume_b200 import sharding
    for count in (0, 5, 64, 6210, 100000):
        for world in (1, 2, 4, 8):
            rs = sharding.shard_ranges(count, world)
            assert rs[0][0] == 0 and rs[-1][1] == count
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 64 + 63


def _shm_worker(rank, world, port, name):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicradiancevolume_b200 import sharding
    H, W = 37, 8
    img = sharding.SharedHostImage(name, (H, W, 4), torch.float16, rank, world)
    band = (H + world - 1) // world
    y0, y1 = min(H, rank * band), min(H, (rank + 1) * band)
    img.tensor[y0:y1] = float(rank + 1)  # every rank writes its own band of rows
    dist.barrier()
    if rank == 0:
        want = torch.zeros(H, W, 4, dtype=torch.float16)
        for r in range(world):
            want[min(H, r * band):min(H, (r + 1) * band)] = float(r + 1)
        assert torch.equal(img.tensor, want)
    dist.barrier()
    img.close()
    dist.destroy_process_group()


def test_shared_host_image_bands_from_every_rank():
    """The end-to-end leg's host image: one POSIX shm segment, every rank fills its band, rank 0 sees them all."""
    mp.spawn(_shm_worker, args=(3, _free_port(), "drv_test_shm_%d" % os.getpid()), nprocs=3, join=True)


def test_interleaved_shards_partition_the_entries():
    """drv_set_shard_interleave: group g of 64 entries belongs to rank g % world; local indices are dense."""
    import ctypes as C
    import dynamicradiancevolume_b200 as drv
    lib = drv.load()
    for count in (0, 1, 63, 64, 65, 6210, 21121):
        for world in (1, 2, 3, 8):
            seen = np.zeros(count, np.int32)
            for rank in range(world):
                n = drv.shard_count(count, rank, world, True)
                for local in range(n):
                    e = C.c_uint32()
                    lib.drv_shard_entry(local, rank, world, C.byref(e))
                    assert e.value < count and (e.value // 64) % world == rank
                    seen[e.value] += 1
            assert np.all(seen == 1)
