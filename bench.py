#!/usr/bin/env python
"""bench.py — the GI-frame benchmark of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            # this repo: libdrv_gi (sm_100a CUDA) through the C-ABI
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the scalar C++ transcription of the
                                                             # reference's GLSL (oracle/), all host threads

metric  = BASELINE.json `metric`: GI ms/frame @1080p, 16k VPLs (configs[1]: procedural atrium 1920x1080, one
          128^2 RSM read, 2 cascades x 64^3, SH1, unshadowed). Lower is better.
step    = one frame of the hot path: RSM mip chain (ShadowMap::PrepareRSM) -> [voxelise + blend + mips when
          indirect shadows are on] -> allocate caches -> VPL generation + cache x VPL gather -> clear HDR ->
          apply (Renderer::Draw, rendering/renderer.cpp:539-594 minus rasterisation / direct light / tonemap).
value   = ms per frame with all inputs resident in HBM, CUDA events on the context's stream around every
          step, L2 flushed (512 MiB memset) between steps, max over ranks.
e2e     = the same frame through the host-buffer C-ABI call (drv_draw_host_frame): pinned host G-buffer + RSM
          level 0 copied H2D and the RGBA16F result copied D2H inside the timed region, copies and stages
          overlapped within the frame (never across frames).
N > 1   = one process per GPU (torchrun). Allocation is replicated (deterministic scan => identical entry
          indices on every rank, no communication); the cache x VPL gather is sharded over contiguous
          cell-ordered entry ranges; finished SH entries are stored to every peer over NVLink from inside the
          gather epilogue (fused all-gather); the cross-GPU barrier is a flag exchange in peer memory
          (drv_peer_barrier; --barrier nccl uses a one-word all-reduce instead); the apply pass is split by pixel
          rows and the RGBA16F bands are gathered on rank 0 with NCCL. One frame is split over N GPUs => "scaling": "strong".
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GI ms/frame @1080p,16k VPLs"
UNIT = "ms/frame"
FLOP_PER_PAIR = {1: 48.0, 2: 92.0}   # 32 ops = 48 flop (SH1), 59 ops = 92 flop (SH2), FMA = 2
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: 148 SMs x 128 lanes x FMA at clocks.max.sm
_MICRO_CACHE = {}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs index (0..3); the metric is quoted on 1. 5 = the reference author's default settings")
    ap.add_argument("--variant", type=int, default=0, help="gather kernel variant (drv_config.gather_variant)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing comparison of the frame with the oracle")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps (profiling runs)")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--stages", action="store_true", help="print the per-stage table to stderr")
    ap.add_argument("--image-gather", choices=["p2p", "nccl"], default="p2p",
                    help="sharded runs: image bands stored into rank 0's target over NVLink by the apply kernel (p2p) or gathered with NCCL")
    ap.add_argument("--no-scaling-workload", action="store_true",
                    help="skip the short run of BASELINE configs[3] (the multi-GPU scaling workload) after the metric's workload")
    ap.add_argument("--static-uniforms", action="store_true",
                    help="do not re-upload PerFrame / VolumeInfo before every frame (a static camera: pure graph replay)")
    ap.add_argument("--no-graph", action="store_true", help="issue the frame kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--serial", action="store_true", help="reference stage order on one stream (prepare_rsm, clear, drv_draw) "
                                                          "instead of drv_draw_frame's light-side || camera-side schedule")
    ap.add_argument("--shard", choices=["interleaved", "contiguous"], default="interleaved",
                    help="sharded runs with the fused NVLink exchange: deal 64-entry groups round-robin to the ranks "
                         "(balances the cone pass) or give every rank one contiguous range of the cell-ordered list")
    ap.add_argument("--barrier", choices=["peer", "nccl"], default="peer",
                    help="cross-GPU barrier of sharded runs: flags in NVLink peer memory (drv_peer_barrier) or an NCCL all-reduce")
    ap.add_argument("--voxel-resolution", type=int, default=0,
                    help="override the workload's voxel volume resolution (shadowed configs; 256 = the record chain no longer fits L2)")
    return ap.parse_args()


def traffic_for(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this command (profiles/ncu_traffic.json names the capture); None if there is no capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def workload_for(index, voxel_resolution=0):
    import workloads
    return workloads.config(index, **({"voxel_resolution": voxel_resolution} if voxel_resolution else {}))


def workload_name(wl):
    return ("%s %dx%d, %d light(s) x %d^2 RSM read (%d VPLs), %d cascade(s) x %d^3, SH%d, %s"
            % (wl.name, wl.width, wl.height, len(wl.lights), int(wl.spot_lights[0].RSMReadResolution), wl.num_vpls,
               wl.cav_cascades, wl.cav_resolution, wl.sh_order,
               "cone-traced indirect shadow, %d^3 voxels" % wl.voxel_resolution if wl.indirect_shadow else "unshadowed"))


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region, sampled in-process through NVML every few
    milliseconds (the timed region of a sub-millisecond frame is too short for `nvidia-smi -lms`, whose first
    sample arrives after ~100 ms); falls back to one `nvidia-smi` query when NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        try:
            power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            power = float("nan")
        self.samples.append((sm, reasons, power))

    def _pump(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is None:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                sm, mx = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": [], "note": "single nvidia-smi query after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML and nvidia-smi unavailable"}
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            try:
                self._sample()
            except Exception:
                pass
        n = self.nvml
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": [], "note": "no samples"}
        bits = 0
        for _, r, _ in self.samples:
            bits |= int(r)
        reasons = sorted(k for k, v in self.REASONS.items() if bits & v)
        return {"sm_mhz": statistics.median([s[0] for s in self.samples]), "sm_max_mhz": mx, "reasons": reasons,
                "power_w_max": max(s[2] for s in self.samples), "samples": len(self.samples),
                "how": "NVML, 2 ms period, over the timed region"}


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's algorithm on the host cores: oracle/ (scalar C++ transcription of the GLSL, std::thread
    over tiles / cache entries). Rank 0 only; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from dynamicradiancevolume_b200 import build as b
    b.build_aux()
    from oracle import binding as orc
    from oracle.frame import OracleFrame
    wl = workload_for(args.config).build()
    cores = orc.default_threads()
    times = []
    for i in range(args.warmup + args.steps):
        o = OracleFrame(wl, threads=cores)
        t = time.perf_counter()
        o.prepare_inputs()
        o.frame()
        dt = (time.perf_counter() - t) * 1e3
        if i >= args.warmup:
            times.append(dt)
    ms = sum(times) / len(times)
    pairs = o.count * wl.num_vpls
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl), "caches": o.count, "vpls": wl.num_vpls},
        "cpu_baseline": {"value": ms, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "whole frame (RSM mips + VPLs + allocate + light + apply), every step, full size",
                         "stage_ms": {k: v * 1e3 for k, v in o.timings.items()},
                         "gather_pairs_per_s": pairs / max(o.timings.get("LightCaches", 0.0), 1e-9)},
        "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- this repo's arm
def run_b200(args):
    """The metric's workload (BASELINE configs[1] unless --config says otherwise) and, beside it, the workload the
    north star quotes multi-GPU scaling on (configs[3]: 3840x2160, 4 lights = 64k VPLs, 4 x 128^3, SH2 + cone-traced
    shadows) as a short device-timed run under "scaling_workload" — the 0.2 ms metric frame is latency-bound and
    cannot strong-scale; the 12 ms one does."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU); WORLD_SIZE is 1" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdrv_gi has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: keep NCCL's own banner / debug output on stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args, args.config, args.steps, args.warmup, light=False)
    if args.config == 1 and not args.no_scaling_workload:
        k = max(3, min(10, args.steps // 3))
        extra = measure(args, 3, k, 2, light=True)
        if line is not None and extra is not None:
            line["scaling_workload"] = {
                "workload": extra["config"]["workload"], "ms_per_frame": extra["value"], "steps": k, "warmup": 2,
                "n_gpus": world, "caches": extra["config"]["caches"], "vpls": extra["config"]["vpls"],
                "live_vpls": extra["config"]["live_vpls"], "scaling": "strong", "parity": extra.get("parity"),
                "stage_ms": extra["stage_ms"], "roofline_cone": extra["roofline_cone"],
                "how": "same timing rules as `value` (CUDA events per step, L2 flushed, max over ranks); "
                       "speed-up at N GPUs = this figure at n_gpus 1 / this figure at N"}
    if args.config == 1 and not args.no_scaling_workload:
        # the same scene with the cone-traced indirect shadows switched on (BASELINE configs[2]): the cone pass is the
        # dominant kernel of every shadowed frame, so its roofline record rides in the default line too
        k = max(3, min(10, args.steps // 3))
        shadowed = measure(args, 2, k, 2, light=True)
        if line is not None and shadowed is not None:
            line["shadowed_workload"] = {
                "workload": shadowed["config"]["workload"], "ms_per_frame": shadowed["value"], "steps": k, "warmup": 2,
                "caches": shadowed["config"]["caches"], "live_vpls": shadowed["config"]["live_vpls"],
                "stage_ms": shadowed["stage_ms"], "roofline_pair_pass": shadowed["roofline"],
                "roofline_cone": shadowed["roofline_cone"], "parity": shadowed.get("parity")}
    if args.config == 1 and not args.no_scaling_workload:
        # the metric's view with a 4x finer address volume (~77 000 caches instead of 6 210): the gather at a cache
        # count where its fixed costs (launch ramp, partial-tile fix-up) no longer dominate
        k = max(3, min(10, args.steps // 3))
        dense = measure(args, 6, k, 2, light=True)
        if line is not None and dense is not None:
            line["dense_workload"] = {
                "workload": dense["config"]["workload"], "ms_per_frame": dense["value"], "steps": k, "warmup": 2,
                "caches": dense["config"]["caches"], "live_vpls": dense["config"]["live_vpls"],
                "stage_ms": dense["stage_ms"], "roofline": dense["roofline"], "parity": dense.get("parity")}
    if line is not None:
        if args.stages:
            for k_, v in line["stage_ms"].items():
                sys.stderr.write("%-18s %8.4f ms\n" % (k_, v))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def measure(args, config_index, n_steps, n_warmup, light):
    """One workload through the C-ABI on this rank's GPU. Returns the JSON line (rank 0) or None. `light`: only
    the device-timed frame (no instrumented pass, no end-to-end leg, no CPU baseline, no micro-benchmarks)."""
    import copy
    args = copy.copy(args)
    args.steps, args.warmup, args.config = n_steps, n_warmup, config_index
    if light:
        args.no_cpu_baseline = args.no_microbench = True
    import numpy as np
    import torch
    import torch.distributed as dist

    import dynamicradiancevolume_b200 as drv
    import workloads
    from dynamicradiancevolume_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    wl = workload_for(args.config, args.voxel_resolution).build()
    stream = torch.cuda.Stream(device=local)
    g = workloads.DeviceFrame(wl, device=local, stream=stream, gather_variant=args.variant)
    ctx = g.ctx
    dev = "cuda:%d" % local
    px = wl.width * wl.height
    # sharded runs: the apply pass is split by rows (sort-first); the bands are gathered on rank 0
    band = (wl.height + world - 1) // world
    hdr16 = torch.zeros(band * world, wl.width, 4, dtype=torch.float16, device=dev)
    band_views = [hdr16[r * band:(r + 1) * band] for r in range(world)]
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    barrier_word = torch.zeros(1, dtype=torch.int32, device=dev)

    if world > 1:  # shard the gather and map every peer's entries buffer (NVLink P2P)
        ctx.set_shard(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.export_entries_ipc())
        for r, h in enumerate(handles):
            if r != rank:
                ctx.import_peer_entries(r, h)
        from dynamicradiancevolume_b200 import sharding
        sharding.connect_image_gather(ctx, rank, world)  # rank 0's RGBA16F target, mapped by every peer
        if args.shard == "interleaved" and args.barrier == "peer" and not args.serial:
            ctx.set_shard_interleave(True)

    def xbarrier():
        if args.barrier == "peer":
            ctx.peer_barrier()
        else:
            dist.all_reduce(barrier_word)

    frame_flags = abi.DRV_FRAME_PREPARE_RSM | (0 if args.no_graph else abi.DRV_FRAME_GRAPH)
    in_frame = not args.serial and (world == 1 or args.barrier == "peer")
    if wl.indirect_shadow and in_frame:
        ctx.bind_scene(g.tris, None, 1.0)  # VoxelizeScene runs inside drv_draw_frame, on its own stream
        frame_flags |= abi.DRV_FRAME_VOXELIZE

    def frame_device():
        with torch.cuda.stream(stream):
            if wl.indirect_shadow and not in_frame:
                ctx.voxelize(g.tris, None, 1.0)
            if not args.static_uniforms:
                # an animated frame: the application uploads PerFrame / VolumeInfo every frame
                # (Renderer::UpdatePerFrameUBO / UpdateVolumeUBO, renderer.cpp:324-431); the recorded frame graph is
                # patched in place with the new kernel arguments (cudaGraphExecUpdate), not re-instantiated
                ctx.set_per_frame(wl.per_frame)
                ctx.set_volume_info(wl.volume)
            if world == 1 and not args.serial:
                # one call: (RSM mips + VPLs) || allocate -> gather -> apply; the glClear of the HDR target
                # (renderer.cpp:562) is fused into the apply pass (DRV_HDR_RGBA16F_WRITE); replayed as a CUDA
                # graph while stage timers are off
                ctx.draw_frame(hdr16, abi.DRV_HDR_RGBA16F_WRITE, frame_flags)
                return
            if world > 1 and not args.serial and args.barrier == "peer":
                # the same call on every rank: allocation replicated, peer barrier, own shard of the gather with
                # the fused all-gather of finished entries, peer barrier, this rank's band of the apply pass;
                # then the image bands are gathered on rank 0
                if args.image_gather == "p2p":
                    # ... and every band is stored straight into rank 0's target over NVLink: no collective at all
                    ctx.draw_frame(None, abi.DRV_HDR_RGBA16F_WRITE, frame_flags | abi.DRV_FRAME_GATHER_IMAGE)
                else:
                    ctx.draw_frame(hdr16, abi.DRV_HDR_RGBA16F_WRITE, frame_flags | abi.DRV_FRAME_APPLY_OWN_ROWS)
                    dist.gather(band_views[rank], band_views if rank == 0 else None, dst=0)
                return
            for i in range(len(g.rsms)):
                ctx.prepare_rsm(i)
            hdr16.zero_()  # glClear(GL_COLOR_BUFFER_BIT), renderer.cpp:562
            if world == 1:
                ctx.draw(hdr16, abi.DRV_HDR_RGBA16F_ADD)
            else:
                ctx.allocate_caches()          # replicated; clears this rank's SH
                xbarrier()                     # every rank has finished clearing before any peer stores arrive
                ctx.light_caches()             # own shard; epilogue stores finished entries to all peers
                xbarrier()                     # all peers' stores have landed
                ctx.apply_caches_rows(hdr16, abi.DRV_HDR_RGBA16F_ADD, rank * band, min(wl.height, (rank + 1) * band))
                dist.gather(band_views[rank], band_views if rank == 0 else None, dst=0)  # the image, on rank 0

    # ---- warm-up + timed region: CUDA events on the context's stream around every step ----
    torch.cuda.synchronize()
    n_total = args.warmup + args.steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_total)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_total)]
    stage_ms = {name: [] for name in abi.STAGE_NAMES}
    step_ms = []
    sampler = ClockSampler(local)
    launches0 = 0
    for i in range(n_total):
        if i == args.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            sampler.start()
            launches0 = ctx.kernel_launches()
            t_wall0 = time.perf_counter()
        with torch.cuda.stream(stream):
            if not args.no_flush:
                flush_buf.zero_()
            if world > 1 and args.barrier == "peer":
                ctx.peer_barrier()  # outside the event pair: every rank starts the step together (no start skew in T)
            ev0[i].record(stream)
        frame_device()
        with torch.cuda.stream(stream):
            ev1[i].record(stream)
        ev1[i].synchronize()
        if i >= args.warmup:
            step_ms.append(ev0[i].elapsed_time(ev1[i]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.kernel_launches() - launches0
    # ---- instrumented pass: the same K steps once more with CUDA events around every stage and around the
    # gather kernel (stage timers force kernel-by-kernel issue, so the frame graph is not used here); the
    # clock sampler keeps running: this pass is part of the measured region of the roofline figures
    ctx.enable_stage_timers(True)
    inst_ms = []
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 4)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 4)]
    n_inst = 3 if light else args.steps + 1
    if wl.indirect_shadow:
        ctx.cone_steps()  # reset the cone pass's sample counter
    for i in range(n_inst):
        with torch.cuda.stream(stream):
            if not args.no_flush:
                flush_buf.zero_()
            ev0[i].record(stream)
        frame_device()
        with torch.cuda.stream(stream):
            ev1[i].record(stream)
        ev1[i].synchronize()
        if i == 0:
            continue
        inst_ms.append(ev0[i].elapsed_time(ev1[i]))
        for s, name in enumerate(abi.STAGE_NAMES):
            try:
                stage_ms[name].append(ctx.stage_ms(s))
            except drv.DrvError:
                pass
    torch.cuda.synchronize()
    cone_steps_per_frame = (ctx.cone_steps() / n_inst) if wl.indirect_shadow else 0
    clocks = sampler.stop()
    ctx.enable_stage_timers(False)
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    n_caches = ctx.active_cache_count()[0]

    # ---- end to end: pinned host inputs -> H2D -> frame -> D2H, through the host-buffer C-ABI calls ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_gb = [pin(a) for a in (wl.depth, wl.normal, wl.diffuse)]
    h_rsm = [[pin(a) for a in r] for r in wl.rsms]
    h_out = torch.zeros(wl.height, wl.width, 4, dtype=torch.float16).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in h_gb) + sum(t.numel() * t.element_size() for r in h_rsm for t in r)
    d2h = h_out.numel() * h_out.element_size()

    # sharded e2e: every rank uploads 1/world of every input over its own PCIe link, an NVLink all-gather completes
    # the images on every GPU (allocation and VPL generation are replicated), then the sharded frame; rank 0 reads
    # the gathered image back
    e2e_sharded = world > 1 and in_frame and args.image_gather == "p2p" and not light
    if e2e_sharded:
        # ONE packed pinned host buffer holds the frame's inputs and ONE device staging buffer receives them: rank r
        # uploads slice r over its own PCIe link, a single NCCL all-gather over NVLink completes the buffer on every
        # GPU. Every rank applies its own band of rows and copies it D2H straight into its rows of a host image that
        # all ranks share (POSIX shared memory, registered with CUDA in every process): no rank-0 funnel.
        host_inputs = list(h_gb) + [t for r in h_rsm for t in r]
        offs, total = [], 0
        for t in host_inputs:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        chunk = ((total + world - 1) // world + 255) // 256 * 256
        host_pack = torch.zeros(chunk * world, dtype=torch.uint8).pin_memory()
        for t, o_ in zip(host_inputs, offs):
            n_ = t.numel() * t.element_size()
            host_pack[o_:o_ + n_].copy_(t.reshape(-1).view(torch.uint8))
        dev_pack = torch.empty(chunk * world, dtype=torch.uint8, device=dev)
        views = [dev_pack[o_:o_ + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for t, o_ in zip(host_inputs, offs)]
        # bound once: the staging images live at fixed addresses, so the recorded frame graph stays valid
        ctx.bind_gbuffer(views[0], views[1], views[2])
        for i in range(len(h_rsm)):
            ctx.bind_rsm(i, views[3 + 3 * i], views[4 + 3 * i], views[5 + 3 * i])
        from dynamicradiancevolume_b200 import sharding
        shared = sharding.SharedHostImage("drv_bench_%s" % os.environ.get("MASTER_PORT", "0"), (wl.height, wl.width, 4),
                                          torch.float16, rank, world)
        h_img = shared.tensor
        y0, y1 = min(wl.height, rank * band), min(wl.height, (rank + 1) * band)
        h2d = chunk
        d2h = (y1 - y0) * wl.width * 8

    def frame_e2e():
        if e2e_sharded:
            with torch.cuda.stream(stream):
                mine = dev_pack[rank * chunk:(rank + 1) * chunk]
                mine.copy_(host_pack[rank * chunk:(rank + 1) * chunk], non_blocking=True)
                dist.all_gather_into_tensor(dev_pack, mine)
                if not args.static_uniforms:
                    ctx.set_per_frame(wl.per_frame)
                    ctx.set_volume_info(wl.volume)
                ctx.draw_frame(hdr16, abi.DRV_HDR_RGBA16F_WRITE, frame_flags | abi.DRV_FRAME_APPLY_OWN_ROWS)
                if y1 > y0:
                    h_img[y0:y1].copy_(hdr16[y0:y1], non_blocking=True)
            stream.synchronize()
            return
        if wl.indirect_shadow:
            ctx.voxelize(g.tris, None, 1.0)
        if world == 1:
            # one call: H2D (RSMs, depth, then normal/albedo bands) -> mips -> allocate -> light -> apply per band
            # -> D2H per band, overlapped inside the frame; returns when the RGBA16F image is in host memory
            ctx.draw_host_frame(h_gb[0], h_gb[1], h_gb[2], h_rsm, h_out)
        else:
            ctx.upload_gbuffer(*h_gb)
            for i, r in enumerate(h_rsm):
                ctx.upload_rsm(i, *r)
                ctx.prepare_rsm(i)
            with torch.cuda.stream(stream):
                hdr16.zero_()
                ctx.allocate_caches()
                xbarrier()
                ctx.light_caches()
                xbarrier()
                ctx.apply_caches_rows(hdr16, abi.DRV_HDR_RGBA16F_ADD, rank * band, min(wl.height, (rank + 1) * band))
                dist.gather(band_views[rank], band_views if rank == 0 else None, dst=0)
                if rank == 0:
                    h_out.copy_(hdr16[:wl.height], non_blocking=True)
            stream.synchronize()

    # what the link gives: one 64 MiB pinned H2D copy and one D2H copy, alone (the floor of the e2e leg is
    # h2d_bytes / this bandwidth: the frame's compute hides behind the input copies)
    pcie = {}
    if not light:
        probe_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        for name, (src, dst) in (("h2d_gbs", (probe_h, probe_d)), ("d2h_gbs", (probe_d, probe_h))):
            best = 1e30
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            pcie[name] = (64 << 20) / best / 1e9
        del probe_h, probe_d
    e2e_ms = []
    for i in range(0 if light else args.warmup + args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        with torch.cuda.stream(stream):
            if not args.no_flush:
                flush_buf.zero_()
        stream.synchronize()
        t0 = time.perf_counter()
        frame_e2e()  # ends with a stream synchronise: the result is in host memory
        dt = (time.perf_counter() - t0) * 1e3
        if i >= args.warmup:
            e2e_ms.append(dt)
    e2e_total = sum(e2e_ms)
    if world > 1:
        t = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_per_step = e2e_total / args.steps
    if e2e_sharded:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:  # the shared host image holds every rank's band: compare it with rank 0's gathered device image later
            e2e_image = h_img.clone()
        shared.close()
    # restore the device-resident bindings (upload_* rebinds to staging copies of the same data)
    ctx.bind_gbuffer(g.depth, g.normal, g.diffuse)
    for i, r in enumerate(g.rsms):
        ctx.bind_rsm(i, *r)

    # ---- parity of what was just timed (after the timed regions; the oracle is the checker, never the thing
    # measured): one more frame through the same call, then rank 0 compares ITS entries / atlas / image — in a
    # sharded run the product of all ranks' shards — with the oracle: allocation bit-exact on the whole frame, SH on
    # every `step`-th entry, the whole image through the oracle's apply pass (oracle/subsample.py)
    parity = None
    if not args.no_parity:
        frame_device()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if rank == 0:
            from oracle.subsample import check_frame
            n_par = ctx.active_cache_count()[0]
            sharded_image = world > 1 and not args.serial and args.barrier == "peer" and args.image_gather == "p2p"
            img_t = ctx.hdr16_tensor() if sharded_image else hdr16
            img = img_t[:wl.height].float().cpu().numpy()
            small = (not wl.indirect_shadow) and n_par * wl.num_vpls <= 2e8
            t0 = time.perf_counter()
            parity = check_frame(wl, ctx.read_entries(n_par), ctx.read_atlas(), n_par, img, step=1 if small else 64,
                                 image_is_half=True)
            parity["seconds"] = time.perf_counter() - t0
            if e2e_sharded:  # the end-to-end leg's host image (bands copied by every rank) against the device frame
                parity["e2e_host_image_equal"] = bool(torch.equal(e2e_image[:wl.height].float(), img_t[:wl.height].float().cpu()))
                parity["ok"] = bool(parity["ok"] and parity["e2e_host_image_equal"])
            parity["what"] = ("allocation (cell set, indices, positions) bit-exact on the whole frame; SH of every %d-th "
                              "entry against the oracle's gather within 1e-5 + 1e-3 rel; the whole RGBA16F image against "
                              "the oracle's apply pass on the device's entries within that gate + one half rounding"
                              % parity["step"])
        if world > 1:
            dist.barrier()

    if rank != 0:
        g.close()
        return None

    # ---- roofline of the dominant kernel (the cache x VPL gather) ----
    med = lambda v: statistics.median(v) if v else None
    avg = lambda v: (sum(v) / len(v)) if v else None
    gather_ms = avg(stage_ms["GatherKernel"])
    cone_ms = avg(stage_ms["ConeKernel"]) if wl.indirect_shadow else None
    if gather_ms and cone_ms:
        gather_ms = max(gather_ms - cone_ms, 1e-6)  # the pair pass alone: GatherKernel brackets cone pass + pair pass
    interleaved = world > 1 and args.shard == "interleaved" and args.barrier == "peer" and not args.serial
    shard_caches = drv.shard_count(n_caches, rank, world, interleaved)
    # units of the roofline: the pairs the kernel EVALUATES (VPLs with zero flux are dropped before the gather;
    # they add exactly zero) — the reference's own pair count (caches x R^2) is reported beside it
    live_vpls = sum(ctx.live_vpl_counts()[:len(wl.spot_lights)])
    pairs_per_launch = shard_caches * live_vpls
    pairs_reference = shard_caches * wl.num_vpls
    flops = pairs_per_launch * FLOP_PER_PAIR[wl.sh_order]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    micro = dict(_MICRO_CACHE)  # the light runs appended to the default line reuse the main run's figures
    if not args.no_microbench:
        lib = drv.load()
        import ctypes as C
        for w in range(lib.drv_microbench_count()):
            name = lib.drv_microbench_name(w).decode()
            if name.startswith("study_"):
                continue  # the operand-delivery study is tools/microbench.py's job
            r = C.c_double()
            if lib.drv_microbench(local, w, C.byref(r)) == 0:
                micro[name] = r.value
        _MICRO_CACHE.update(micro)
    fp32_peak = micro.get("ffma_tflops") or NOMINAL_FP32_TFLOPS
    roofline = None
    if gather_ms:
        achieved = flops / (gather_ms * 1e-3) / 1e12
        roofline = {
            "kernel": "gather_ws_kernel<SH%d,%s>" % (wl.sh_order, "shadow" if wl.indirect_shadow else "unshadowed"),
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "peak_source": ("measured on this GPU at bench start: scalar-FFMA micro-kernel (drv_microbench), FMA = 2 flop"
                            if "ffma_tflops" in micro else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"),
            "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
            "traffic": traffic_for("gather_ws_kernel<SH%d,%s>" % (wl.sh_order, "shadow" if wl.indirect_shadow else "unshadowed")),
            "pairs_per_launch": pairs_per_launch, "flop_per_pair": FLOP_PER_PAIR[wl.sh_order],
            "live_vpls": live_vpls, "pairs_per_launch_reference": pairs_reference,
            "frac_counting_reference_pairs": pairs_reference * FLOP_PER_PAIR[wl.sh_order] / (gather_ms * 1e-3) / 1e12 / fp32_peak,
            "pairs_per_s": pairs_per_launch / (gather_ms * 1e-3), "avg_launch_ms": gather_ms,
            "note": ("CUDA-core FP32 roofline (this is not a tensor-core contraction; MEASURED_PEAKS.json has no FP32 "
                     "figure). HBM traffic of the gather is ~0 per pair: the VPL list and entries are L2-resident."),
        }
    # ---- the cone pass (the dominant kernel of the shadowed workloads): SURVEY 8d row "cone trace" ----
    roofline_cone = None
    if cone_ms and cone_steps_per_frame:
        steps_per_s = cone_steps_per_frame / (cone_ms * 1e-3)
        lane_ops_peak = (micro.get("ffma_tflops") or NOMINAL_FP32_TFLOPS) * 1e12 / 2.0  # lane-ops/s (FMA = 2 flop)
        l2_peak = micro.get("l2_read_gbs")
        texel_gbs = steps_per_s * 16.0 / 1e9
        roofline_cone = {
            "kernel": "cone_kernel", "unit": "cone steps (voxel samples)", "steps_per_launch": cone_steps_per_frame,
            "avg_launch_ms": cone_ms, "steps_per_s": steps_per_s,
            "texel_bytes_per_step": 16, "texel_gbs": texel_gbs, "l2_read_gbs_measured": l2_peak,
            "frac_of_l2": (texel_gbs / l2_peak) if l2_peak else None,
            "fp32_ops_per_step": 45, "fp32_frac": steps_per_s * 45.0 / lane_ops_peak,
            "bound": "issue (instruction count per sample), not bandwidth: see profiles/ ncu summary",
            "note": "algorithmic figures of SURVEY 8d: 16 UNORM8 texels (2 mips x 8) and ~45 FP32 ops per step; the kernel "
                    "reads one 8-byte record per mip level per step from the L2-resident record chain"}
    # HBM-side summary of the two streaming stages (algorithmic bytes)
    stride = abi.entry_stride(wl.sh_order)
    cells = wl.cav_cascades * wl.cav_resolution ** 3
    alloc_bytes = 4 * px + cells * (1 + 1 + 1 + 4) + n_caches * stride
    apply_bytes = px * 24  # depth 4 + normal 4 + diffuse 4 (sRGB8 padded) + RGBA16F 8 (+ 4 of slack it allows)
    hbm_peak = peaks.get("hbm_gbs")
    secondary = {}
    for name, nbytes in (("AllocateCaches", alloc_bytes), ("ApplyCaches", apply_bytes)):
        m = med(stage_ms[name])
        if m:
            gbs = nbytes / (m * 1e-3) / 1e9
            secondary[name] = {"bound": "hbm", "algorithmic_bytes": nbytes, "ms": m, "achieved": gbs, "unit": "GB/s",
                               "peak": hbm_peak, "frac": (gbs / hbm_peak) if hbm_peak else None}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import binding as orc
        from oracle.frame import OracleFrame
        cores = orc.default_threads()
        runs = []
        for _ in range(3):
            o = OracleFrame(wl, threads=cores)
            t0 = time.perf_counter()
            o.prepare_inputs()
            o.frame()
            runs.append((time.perf_counter() - t0) * 1e3)
        cpu_baseline = {"value": statistics.median(runs), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "the whole frame at full size, median of 3 runs (oracle/: scalar C++ transcription "
                                  "of the reference GLSL, std::thread over tiles / cache entries)",
                        "gather_pairs_per_s": o.count * wl.num_vpls / max(o.timings.get("LightCaches", 0.0), 1e-9)}

    line = {
        "metric": METRIC, "value": ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl), "caches": n_caches, "vpls": wl.num_vpls, "live_vpls": live_vpls,
                   "pairs_per_frame": n_caches * wl.num_vpls,
                   "l2": "flushed between steps (512 MiB memset outside the event pairs)" if not args.no_flush else "not flushed",
                   "parallelism": "1 GPU" if world == 1 else (
                       "gather sharded over %d GPUs by %s, allocation replicated, fused P2P all-gather of SH, apply "
                       "row-sharded, image bands %s" % (
                           world,
                           "64-entry groups of the cell-ordered list dealt round-robin" if interleaved else "cell-ordered entry range",
                           "stored into rank 0's target over NVLink (no collective in the frame)"
                           if (args.image_gather == "p2p" and args.barrier == "peer" and not args.serial) else "gathered on rank 0 with NCCL")),
                   "gather_variant": args.variant},
        "e2e": {"value": e2e_per_step, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "link": pcie,
                "h2d_floor_ms": (h2d / (pcie["h2d_gbs"] * 1e9) * 1e3) if pcie.get("h2d_gbs") else None,
                "how": ("drv_draw_host_frame: pinned host G-buffer + RSM level 0 -> H2D -> mips, allocate, light, apply per "
                        "band -> D2H per band, overlapped inside the frame; wall clock around the call, which returns when "
                        "the RGBA16F image is in host memory") if world == 1 else
                       ("every rank uploads slice 1/%d of ONE packed pinned input buffer over its own PCIe link, a single NCCL "
                        "all-gather over NVLink completes it on every GPU, sharded drv_draw_frame with each rank applying its "
                        "own band of rows, each rank copies its band D2H straight into a host image shared by all ranks "
                        "(POSIX shm registered with CUDA); h2d / d2h bytes are PER RANK; wall clock, max over ranks" % world if e2e_sharded else
                        "every rank uploads all inputs, serial sharded stages, NCCL image gather, rank 0 D2H")},
        "gpu_launches": launches,
        "parity": parity,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_cone": roofline_cone,
        "roofline_streaming_stages": secondary,
        "cpu_baseline": cpu_baseline,
        "stage_ms": {k: med(v) for k, v in stage_ms.items() if v},
        "stage_ms_note": "instrumented pass after the timed region: the same steps issued kernel by kernel with CUDA "
                         "events around every stage (%.4f ms/frame that way); stages of the light side and the camera "
                         "side overlap" % (sum(inst_ms) / max(len(inst_ms), 1)),
        "frame_issue": ("serial: prepare_rsm, clear, drv_draw" if (args.serial or (world > 1 and args.barrier != "peer")) else
                        "drv_draw_frame: (RSM mips + VPLs) || [voxelise] || allocate -> gather -> apply(+clear)%s"
                        % ("" if args.no_graph else (", CUDA graph replay" if args.static_uniforms else
                                                     ", CUDA graph patched with the frame's uniforms (cudaGraphExecUpdate) and replayed"))),
        "graph": dict(zip(("instantiations", "updates"), ctx.graph_stats())),
        "microbench": micro,
        "wall_ms_per_step_incl_flush": t_wall * 1e3 / args.steps,
        "gpu": torch.cuda.get_device_name(local),
    }
    g.close()
    return line


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that write banners to file descriptor 1 (NCCL's version
    # line, for one) are sent to stderr for the duration of the run, and print() gets the real stdout back
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        rc = run_reference(args)
    else:
        rc = run_b200(args)
    sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
