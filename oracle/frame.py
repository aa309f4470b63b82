"""Whole-frame driver of the CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Runs the reference's stage order (Renderer::Draw, rendering/renderer.cpp:539-594)
over a ``workloads.Workload`` with the scalar restatement of the shaders:
RSM mips -> VPLs (+ shadow blocks) -> voxelise/blend/mips -> allocate -> light -> apply.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import time

import numpy as np

from . import binding as orc


class OracleFrame:
    def __init__(self, wl, threads: int = 0):
        self.wl = wl
        self.threads = threads
        self.timings = {}
        self.levels = []
        self.vpls = []
        self.blocks = []
        self.chain = None
        self.alloc = None
        self.entries = None
        self.image = None

    def _timed(self, name, fn):
        t = time.perf_counter()
        r = fn()
        self.timings[name] = self.timings.get(name, 0.0) + (time.perf_counter() - t)
        return r

    def prepare_inputs(self):
        """downsamplersm.frag mip chains, the VPL lists / shadow-block records, and the voxel chain."""
        import workloads
        wl = self.wl

        def rsm():
            self.levels = [orc.rsm_mip_chain(*r) for r in wl.rsms]
        self._timed("PrepareRSM", rsm)

        def vpl():
            self.vpls, self.blocks = [], []
            for s, lv in zip(wl.spot_lights, self.levels):
                rl = workloads.rsm_read_level(s)
                self.vpls.append(orc.generate_vpls(s, *lv[rl]))
                if wl.indirect_shadow:
                    self.blocks.append(orc.shadow_blocks(s, lv[rl + int(s.IndirectShadowComputationLod)][2]))
                else:
                    self.blocks.append(None)
        self._timed("GenerateVPLs", vpl)
        if wl.indirect_shadow:
            def vox():
                res = wl.voxel_resolution
                target = orc.voxelize(wl.volume, res, wl.triangles)
                vol = np.zeros(res ** 3, np.uint8)
                orc.voxel_blend(vol, target, res, 1.0)
                self.target = target
                self.chain = orc.voxel_chain(vol, res)
            self._timed("VoxelizeScene", vox)
        return self

    def allocate(self):
        wl = self.wl
        self.alloc = self._timed("AllocateCaches", lambda: orc.allocate_caches(
            wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.depth, wl.sh_order, wl.max_caches, self.threads))
        self.count = self.alloc["count"]
        self.entries = self.alloc["entries"]
        return self

    def light(self, first=0, count=None, fp64=False, entries=None):
        wl = self.wl
        e = self.entries if entries is None else entries
        n = self.count - first if count is None else count
        self._timed("LightCaches", lambda: orc.light_caches(
            wl.constant, wl.volume, wl.spot_lights, self.vpls, self.blocks, self.chain, e, first, n, wl.sh_order,
            wl.indirect_shadow, fp64, self.threads))
        return e

    def apply(self, entries=None):
        wl = self.wl
        e = self.entries if entries is None else entries
        self.image = self._timed("ApplyCaches", lambda: orc.apply_caches(
            wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal, wl.diffuse,
            self.alloc["atlas"], e, self.threads))
        return self.image

    def frame(self):
        self.allocate()
        self.light()
        return self.apply()


def close(a, b, rtol=1e-3, atol=1e-5):
    """The north-star gate: |a-b| <= atol + rtol * max(|a|,|b|). Returns (ok, worst excess ratio)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.maximum(np.abs(a), np.abs(b))
    ratio = float(np.max(err / tol)) if err.size else 0.0
    return bool(np.all(err <= tol)), ratio
