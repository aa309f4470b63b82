#!/usr/bin/env python
"""Device timings of the rows next to the hot path on the metric's 1080p scene: RSM fill (1024^2),
voxel cone-traced AO (128^3 volume), tonemap — CUDA events, median of `reps`, with the HBM roofline of the two
streaming kernels (algorithmic bytes / time against MEASURED_PEAKS.json).

    python tools/adjacent_bench.py
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import workloads
    wl = workloads.config(2).build()  # 1080p, 128^3 voxels
    stream = torch.cuda.Stream()
    g = workloads.DeviceFrame(wl, device=0, stream=stream)
    ctx = g.ctx
    with torch.cuda.stream(stream):
        g.prepare_inputs()
    R = int(wl.spot_lights[0].RSMRenderResolution)
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(rng.uniform(-5, 5, size=(R, R, 3)).astype(np.float32)).cuda()
    nrm = torch.from_numpy(rng.normal(size=(R, R, 3)).astype(np.float32)).cuda()
    base = torch.from_numpy(rng.uniform(0, 1, size=(R, R, 3)).astype(np.float32)).cuda()
    ao = torch.zeros(wl.height, wl.width, dtype=torch.float32, device="cuda")
    hdr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16, device="cuda")
    ldr = torch.zeros(wl.height, wl.width, 4, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs")

    def timed(fn, reps=20):
        ts = []
        for i in range(reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                fn()
                e1.record(stream)
            e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    px = wl.width * wl.height
    out = {}
    ms = timed(lambda: ctx.fill_rsm(0, pos, nrm, base, None))
    b = R * R * (36 + 16)
    out["fill_rsm"] = {"ms": ms, "texels": R * R, "algorithmic_bytes": b, "gbs": b / ms / 1e6, "hbm_frac": (b / ms / 1e6 / hbm) if hbm else None}
    ms = timed(lambda: ctx.cone_trace_ao(ao))
    out["cone_trace_ao"] = {"ms": ms, "pixels": px, "cones": 6 * px, "max_steps_per_cone": 16,
                            "note": "issue / L1-L2 gather bound like the cone pass of the gather; not a streaming kernel"}
    ms = timed(lambda: ctx.tonemap(hdr, ldr, 1.0, 1.2))
    b = px * (8 + 16)
    out["tonemap"] = {"ms": ms, "pixels": px, "algorithmic_bytes": b, "gbs": b / ms / 1e6, "hbm_frac": (b / ms / 1e6 / hbm) if hbm else None}
    # the oracle (scalar C++ restatement, all host cores for AO; fill / tonemap are single-threaded loops) beside it
    import time
    from oracle import binding as orc
    from oracle.frame import OracleFrame
    o = OracleFrame(wl).prepare_inputs()
    t0 = time.perf_counter()
    orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal)
    out["cone_trace_ao"]["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    out["cone_trace_ao"]["cpu_cores"] = orc.default_threads()
    t0 = time.perf_counter()
    orc.fill_rsm(wl.spot_lights[0], pos.cpu().numpy(), nrm.cpu().numpy(), base.cpu().numpy())
    out["fill_rsm"]["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    out["fill_rsm"]["cpu_cores"] = 1
    h = np.zeros((wl.height, wl.width, 4), np.float32)
    t0 = time.perf_counter()
    orc.tonemap(h, 1.0, 1.0)
    out["tonemap"]["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    out["tonemap"]["cpu_cores"] = 1
    print(json.dumps(out))


if __name__ == "__main__":
    main()
