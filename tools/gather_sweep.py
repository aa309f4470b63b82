#!/usr/bin/env python
"""BASELINE config 5: the gather sweep — synthetic cache entries x VPLs (no scene) through the C-ABI, timing the
cache x VPL kernel alone (DRV_STAGE_GATHER_KERNEL CUDA events) per kernel variant. One JSON line per point.

    python tools/gather_sweep.py [--caches 65536,1048576] [--vpls 4096,16384] [--orders 1,2] [--variants 0,2,3,4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP = {1: 48.0, 2: 92.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--caches", default="65536,262144,1048576")
    ap.add_argument("--vpls", default="4096,16384,65536")
    ap.add_argument("--orders", default="1,2")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--budget", type=float, default=4e12, help="skip points with more pairs than this")
    a = ap.parse_args()
    import torch
    import dynamicradiancevolume_b200 as drv
    import workloads
    ints = lambda s: [int(x) for x in s.split(",") if x]
    for order in ints(a.orders):
        for n_cache in ints(a.caches):
            for n_vpl in ints(a.vpls):
                if float(n_cache) * n_vpl > a.budget:
                    continue
                pos, vpls = workloads.sweep(n_cache, n_vpl)
                p = torch.from_numpy(pos).cuda()
                for variant in ints(a.variants):
                    ctx = drv.Context(max_cache_count=n_cache, cav_cascades=1, cav_resolution=8, voxel_resolution=16,
                                      sh_order=order, indirect_shadow=False, cascade_transitions=False, width=16,
                                      height=16, max_lights=1, max_rsm_resolution=1 << (max(n_vpl - 1, 1).bit_length() + 1) // 2,
                                      gather_variant=variant)
                    ctx.set_constant(drv.pack_constant(16, 16, 16, 8, 1, n_cache))
                    ctx.set_light_count(1)
                    torch.cuda.synchronize()
                    ctx.set_synthetic_entries(p)
                    ctx.set_vpls(0, vpls.ctypes.data, n_vpl)
                    ctx.enable_stage_timers(True)
                    ms = []
                    for r in range(a.reps + 2):
                        ctx.light_caches()
                        t = ctx.stage_ms(6)
                        if r >= 2:
                            ms.append(t)
                    best = min(ms)
                    pairs = float(n_cache) * n_vpl
                    print(json.dumps({"sh_order": order, "caches": n_cache, "vpls": n_vpl, "variant": variant,
                                      "gather_ms": best, "pairs_per_s": pairs / (best * 1e-3),
                                      "tflops": pairs * FLOP[order] / (best * 1e-3) / 1e12}), flush=True)
                    ctx.close()


if __name__ == "__main__":
    main()
