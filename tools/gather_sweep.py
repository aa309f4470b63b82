#!/usr/bin/env python
"""BASELINE configs[4]: the gather sweep — synthetic cache entries x VPLs (no scene; SURVEY 8d "C5") through the
C-ABI, timing the cache x VPL kernel alone (DRV_STAGE_GATHER_KERNEL CUDA events) per kernel variant, at 1 GPU or —
under torchrun — sharded over N GPUs with the fused NVLink exchange of finished entries. One JSON line per point:
median (and min) of `--reps` launches after warm-up, every launch behind an L2 flush (which also keeps the host
ahead of the device, so launch latency is not part of the figure), pairs/s, TFLOP/s and fraction of the FP32
micro-benchmark, the SM clock / throttle record over the point's timed region, and the worst |err|/tol of a stated
subsample of entries against the CPU oracle (`--check` entries, evenly spaced).

    python tools/gather_sweep.py [--caches 65536,1048576,4194304] [--vpls 4096,16384,65536,262144] [--orders 1,2]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/gather_sweep.py ...
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP = {1: 48.0, 2: 92.0}


def oracle_subsample(cb, pos, vpls, sh_order, idx):
    """Oracle SH of the entries `idx` (compact copy of their positions), float accumulation."""
    from dynamicradiancevolume_b200 import abi
    from oracle import binding as orc
    stride = abi.entry_stride(sh_order) // 4
    e = np.zeros((len(idx), stride), np.float32)
    e[:, :3] = pos[idx, :3]
    s = abi.SpotLight()
    s.RSMReadResolution = 1  # the oracle walks RSMReadResolution^2 VPLs: give it the list in rows of one
    # orc.light_caches needs a square count; feed the list in square chunks and accumulate
    n = len(vpls)
    done = 0
    while done < n:
        r = int(np.floor(np.sqrt(n - done)))
        s.RSMReadResolution = r
        orc.light_caches(cb, abi.VolumeInfo(), [s], [np.ascontiguousarray(vpls[done:done + r * r])], None, None, e, 0,
                         len(idx), sh_order, False)
        done += r * r
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--caches", default="65536,262144,1048576,4194304")
    ap.add_argument("--vpls", default="4096,16384,65536,262144")
    ap.add_argument("--orders", default="1,2")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--budget", type=float, default=1.2e12, help="skip points with more pairs PER GPU than this")
    ap.add_argument("--check", type=int, default=48, help="entries compared with the CPU oracle per point (0 = none)")
    ap.add_argument("--check-pairs", type=float, default=3e8, help="cap of oracle pairs per point (entries are reduced to fit)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dynamicradiancevolume_b200 as drv
    import workloads
    from bench import ClockSampler
    from oracle.frame import close
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from dynamicradiancevolume_b200 import sharding
    lib = drv.load()
    import ctypes as C
    peak = C.c_double()
    lib.drv_microbench(local, 0, C.byref(peak))  # scalar-FFMA TFLOP/s on this GPU
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ints = lambda s: [int(x) for x in s.split(",") if x]
    for order in ints(a.orders):
        for n_cache in ints(a.caches):
            for n_vpl in ints(a.vpls):
                if float(n_cache) * n_vpl / world > a.budget:
                    continue
                pos, vpls = workloads.sweep(n_cache, n_vpl)
                p = torch.from_numpy(pos).to(dev)
                for variant in ints(a.variants):
                    stream = torch.cuda.Stream(device=local)
                    rsm_cap = 1 << (max(n_vpl - 1, 1).bit_length() + 1) // 2
                    ctx = drv.Context(max_cache_count=n_cache, cav_cascades=1, cav_resolution=8, voxel_resolution=16,
                                      sh_order=order, indirect_shadow=False, cascade_transitions=False, width=16,
                                      height=16, max_lights=1, max_rsm_resolution=rsm_cap, gather_variant=variant,
                                      device=local, stream=stream.cuda_stream)
                    cb = drv.pack_constant(16, 16, 16, 8, 1, n_cache)
                    ctx.set_constant(cb)
                    ctx.set_light_count(1)
                    if world > 1:
                        sharding.connect_peers(ctx, rank, world)
                        ctx.set_shard_interleave(True)
                    torch.cuda.synchronize()
                    ctx.set_vpls(0, vpls.ctypes.data, n_vpl)
                    ctx.enable_stage_timers(True)
                    ms = []
                    sampler = ClockSampler(local)
                    for r in range(a.reps + 2):
                        if r == 2:
                            sampler.start()
                        with torch.cuda.stream(stream):
                            ctx.set_synthetic_entries(p)  # SH back to zero
                            flush.zero_()
                            if world > 1:
                                ctx.peer_barrier()
                            ctx.light_caches()
                            if world > 1:
                                ctx.peer_barrier()  # every peer's stores have landed
                        stream.synchronize()
                        if r >= 2:
                            ms.append(ctx.stage_ms(6))
                    clocks = sampler.stop()
                    med, best = statistics.median(ms), min(ms)
                    if world > 1:  # the slowest rank defines the point
                        t = torch.tensor([med, best], dtype=torch.float64, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        med, best = float(t[0]), float(t[1])
                    parity = None
                    if a.check and rank == 0:
                        k = max(1, min(a.check, int(a.check_pairs // n_vpl)))
                        idx = np.unique(np.linspace(0, n_cache - 1, k).astype(np.int64))
                        e = ctx.read_entries(n_cache)[idx]  # after the exchange: entries of every rank's shard
                        eo = oracle_subsample(cb, pos, vpls, order, idx)
                        ok, ratio = close(e[:, 4:], eo[:, 4:])
                        parity = {"checked_entries": int(len(idx)), "worst_ratio": ratio, "ok": bool(ok)}
                    pairs = float(n_cache) * n_vpl
                    if rank == 0:
                        tf = pairs * FLOP[order] / (med * 1e-3) / 1e12
                        print(json.dumps({"sh_order": order, "caches": n_cache, "vpls": n_vpl, "variant": variant,
                                          "n_gpus": world, "gather_ms_median": med, "gather_ms_min": best, "reps": a.reps,
                                          "pairs_per_s": pairs / (med * 1e-3), "tflops": tf,
                                          "frac_of_fp32_peak_all_gpus": tf / (peak.value * world),
                                          "fp32_peak_tflops_per_gpu": peak.value, "clocks": clocks, "parity": parity,
                                          "timing": "CUDA events around the kernel, L2 flushed before every launch, "
                                                    "median of reps; N > 1: peer barrier before and after, max over ranks"}),
                              flush=True)
                    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
