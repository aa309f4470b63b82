#!/usr/bin/env python
"""End-to-end frame (drv_draw_host_frame) wall time for several band counts, next to the raw copy times of the
same buffers — where the e2e leg of bench.py spends its time.

    python tools/e2e_probe.py [--config 1]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import numpy as np
    import torch
    import workloads
    wl = workloads.config(a.config).build()
    stream = torch.cuda.Stream()
    g = workloads.DeviceFrame(wl, device=0, stream=stream)
    ctx = g.ctx
    ctx.enable_stage_timers(True)  # before the first host frame: its events then carry time stamps (the timeline)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    h_gb = [pin(x) for x in (wl.depth, wl.normal, wl.diffuse)]
    h_rsm = [[pin(x) for x in r] for r in wl.rsms]
    h_out = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16).pin_memory()
    if wl.indirect_shadow:
        ctx.voxelize(g.tris, None, 1.0)
    out = {}
    for bands in (1, 2, 4, 8, 16, 32):
        ts = []
        for i in range(a.reps + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.draw_host_frame(h_gb[0], h_gb[1], h_gb[2], h_rsm, h_out, bands)
            dt = (time.perf_counter() - t0) * 1e3
            if i >= 3:
                ts.append(dt)
        out["bands_%d_ms" % bands] = sorted(ts)[len(ts) // 2]
        if bands in (1, 4):
            tl = ctx.host_frame_timeline()
            out["timeline_bands_%d" % bands] = {k: (round(v, 4) if not isinstance(v, list) else [[round(x, 4) for x in b] for b in v])
                                                for k, v in tl.items()}
    # raw copies of the same buffers, back to back on one stream
    dev = [torch.empty_like(t, device="cuda") for t in h_gb] + [torch.empty_like(t, device="cuda") for r in h_rsm for t in r]
    src = h_gb + [t for r in h_rsm for t in r]
    d_out = torch.zeros_like(h_out, device="cuda")
    ts, ts2 = [], []
    for i in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s, d in zip(src, dev):
            d.copy_(s, non_blocking=True)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ts2.append((time.perf_counter() - t0) * 1e3)
    out["h2d_all_inputs_ms"] = sorted(ts)[len(ts) // 2]
    out["d2h_image_ms"] = sorted(ts2)[len(ts2) // 2]
    out["h2d_bytes"] = sum(t.numel() * t.element_size() for t in src)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
