#!/usr/bin/env python
"""Where the fixed cost of the cache x VPL kernel goes: runs the metric's frame (BASELINE configs[1]) with
gather_variant bit 18 set, reads the per-CTA %globaltimer stamps of the pair kernel and prints, relative to the
earliest CTA entry, when CTAs enter, finish their prologue / pair loop, and leave.

    python tools/gather_trace.py [--variant 0] [--config 1] [--frames 6]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fixup(t, t0):
    """The CTAs that drew a tile's last ticket: (ticket drawn, partial sums visible, entries written), relative."""
    last = t[t[:, 7] > t0]
    if len(last) == 0:
        return None
    return [[float((r[5] - t0) / 1e3), float((r[6] - t0) / 1e3), float((r[7] - t0) / 1e3), float((r[3] - t0) / 1e3)] for r in last[:6]]


def _pair_gap(t):
    by_sm = {}
    for row in t:
        by_sm.setdefault(int(row[4]), []).append(row[2])
    gaps = [(max(v) - min(v)) / 1e3 for v in by_sm.values() if len(v) > 1]
    return [float(np.min(gaps)), float(np.median(gaps)), float(np.max(gaps))] if gaps else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--shard-world", type=int, default=1, help="light only shard 0 of this many (no peers: what one rank of a sharded run computes)")
    a = ap.parse_args()
    import torch
    import workloads
    from dynamicradiancevolume_b200 import abi
    wl = workloads.config(a.config).build()
    stream = torch.cuda.Stream()
    g = workloads.DeviceFrame(wl, device=0, stream=stream, gather_variant=a.variant | 0x40000)
    if a.shard_world > 1:
        g.ctx.set_shard(0, a.shard_world)
    out16 = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    if wl.indirect_shadow:
        g.ctx.bind_scene(g.tris, None, 1.0)
    flags = abi.DRV_FRAME_PREPARE_RSM | abi.DRV_FRAME_GRAPH | (abi.DRV_FRAME_VOXELIZE if wl.indirect_shadow else 0)
    rows = []
    for i in range(a.frames):
        with torch.cuda.stream(stream):
            flush.zero_()
            g.ctx.draw_frame(out16, abi.DRV_HDR_RGBA16F_WRITE, flags)
        stream.synchronize()
        t = g.ctx.gather_trace().astype(np.int64)
        t = t[t[:, 0] > 0]
        if i < 2 or len(t) == 0:
            continue
        t0 = t[:, 0].min()
        r = (t - t0) / 1e3
        rows.append({"ctas": int(len(t)),
                     "enter_us": [float(r[:, 0].min()), float(np.median(r[:, 0])), float(r[:, 0].max())],
                     "p1_us": [float(r[:, 1].min()), float(np.median(r[:, 1])), float(r[:, 1].max())],
                     "p2_us": [float(r[:, 2].min()), float(np.median(r[:, 2])), float(r[:, 2].max())],
                     "exit_us": [float(r[:, 3].min()), float(np.median(r[:, 3])), float(r[:, 3].max())],
                     "cta_busy_us": [float((r[:, 3] - r[:, 0]).min()), float(np.median(r[:, 3] - r[:, 0])),
                                     float((r[:, 3] - r[:, 0]).max())],
                     "fixup_us": _fixup(t, t0),
                     "sms_used": int(len(set(t[:, 4].tolist()))),
                     "ctas_per_sm_max": int(np.bincount(t[:, 4].astype(np.int64)).max()),
                     # the two CTAs of an SM: how far apart they finish their pair loops
                     "pair_gap_us": _pair_gap(t)})
    for r in rows:
        print(json.dumps({"variant": a.variant, "config": a.config, "shard_world": a.shard_world, **r}))
    g.close()


if __name__ == "__main__":
    main()
