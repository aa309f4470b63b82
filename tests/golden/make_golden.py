#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors for this path and cannot be built or run in this environment (GLSL 4.50
on a Win32 / OpenGL 4.5 host), so these fixtures are REGRESSION pins of the oracle itself: they freeze
its outputs (checked by the hand-derived / numpy cross-checks of tests/test_oracle_kat.py at the time they were
made) so later edits to oracle/ or scenes/ cannot drift unnoticed, and they give the GPU tests a fixed target
that does not depend on the oracle being rebuilt on the GPU box.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    import workloads
    return {
        "cornell_sh2_shadow": workloads.cornell(width=96, height=96, rsm_res=64, read_lod=1, sh_order=2,
                                                indirect_shadow=True, cav_resolution=16, voxel_resolution=32,
                                                shadow_lod=1),
        "atrium_sh1_transitions": workloads.atrium(width=160, height=90, rsm_res=32, read_lod=0, sh_order=1,
                                                   cav_resolution=16, first_cascade=8.0, max_caches=8192),
    }


def input_digest(wl):
    h = hashlib.sha256()
    for a in [wl.depth, wl.normal, wl.diffuse, wl.triangles] + [x for r in wl.rsms for x in r]:
        h.update(np.ascontiguousarray(a).tobytes())
    for b in [wl.constant, wl.per_frame, wl.volume] + list(wl.spot_lights):
        h.update(bytes(b))
    return h.hexdigest()


def run(wl):
    from oracle import binding as orc
    from oracle.frame import OracleFrame
    wl.build()
    o = OracleFrame(wl).prepare_inputs()
    img = o.frame()
    ids = orc.allocated_cell_ids(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.depth)
    out = dict(digest=np.array(input_digest(wl)), cell_ids=ids, entries=o.entries[:o.count].copy(),
               image=img[..., :3].copy(), shaded=np.packbits(img[..., 3] > 0))
    if wl.indirect_shadow:
        out["voxel_chain_sha256"] = np.array(hashlib.sha256(o.chain.tobytes()).hexdigest())
        out["voxel_set"] = np.packbits(o.target > 0)
        out["shadow_blocks"] = o.blocks[0].view(np.float32).reshape(-1, 4).copy()
    out["vpl_head"] = o.vpls[0][:64].view(np.float32).reshape(-1, 12).copy()
    # rows next to the path: AO through the same voxel chain, tonemap of the frame, RSM fill of seeded
    # fragment attributes
    if wl.indirect_shadow:
        out["ao"] = orc.cone_trace_ao(wl.per_frame, wl.volume, o.chain, wl.voxel_resolution, wl.depth, wl.normal)
    out["tonemap"] = orc.tonemap(img, 2.0, np.float32(np.log2(2.2)))
    pos, nrm, base, cov = fill_rsm_inputs(wl)
    fo, no, do = orc.fill_rsm(wl.spot_lights[0], pos, nrm, base, cov)
    out["fill_rsm_flux"], out["fill_rsm_normal"], out["fill_rsm_depth"] = fo, no, do
    return out


def fill_rsm_inputs(wl):
    """Seeded per-fragment attributes for light 0 (32 x 32 of its render resolution's pixel solid angle)."""
    L = wl.spot_lights[0]
    R = 32
    rng = np.random.default_rng(0xF111)
    lp = np.array(L.LightPosition[:3], np.float32)
    ld = np.array(L.LightDirection[:3], np.float32)
    d = ld + rng.normal(size=(R, R, 3)).astype(np.float32) * 0.5
    pos = (lp + d * rng.uniform(0.5, 6.0, size=(R, R, 1)).astype(np.float32)).astype(np.float32)
    nrm = rng.normal(size=(R, R, 3)).astype(np.float32)
    base = rng.uniform(0, 1, size=(R, R, 3)).astype(np.float32)
    cov = (rng.uniform(size=(R, R)) > 0.2).astype(np.uint8)
    return pos, nrm, base, cov


if __name__ == "__main__":
    for name, wl in cases().items():
        out = run(wl)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "caches", len(out["cell_ids"]), os.path.getsize(path), "bytes")
