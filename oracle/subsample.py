"""Stated-subsample parity of a FULL-SIZE frame — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

The oracle cannot light every cache of BASELINE configs[3] (21 k caches x 65 k VPLs with 4096 cone marches each) in
seconds, so full-size frames are checked the way BASELINE.md section 3 states:
  * allocation: the whole frame — cell set, atlas indices and entry positions bit-exact;
  * cache x VPL gather (+ cone-traced visibility): every `step`-th entry, all lights, the full VPL lists, within the
    north-star gate |a-b| <= 1e-5 + 1e-3 max(|a|,|b|);
  * apply: the whole image, computed by the oracle FROM THE ENTRIES UNDER TEST (the SH the device produced, of which
    the subsample has just been verified), within the same gate — or, for an RGBA16F target, within the gate plus
    one half-precision rounding.
Used by tests/test_gpu_frame.py (full-size configs[3]) and by bench.py's post-timing parity leg at every GPU count.
"""
import numpy as np

from . import binding as orc
from .frame import OracleFrame, close


def check_frame(wl, entries, atlas, count, image=None, step=64, image_is_half=False, threads=0, oracle=None):
    """entries [>=count, stride/4] float32, atlas [R, R, R*C] uint32, image [H, W, >=3] float — what the device
    produced for workload `wl`. Returns a dict; `ok` is the conjunction of every gate."""
    o = oracle if oracle is not None else OracleFrame(wl, threads=threads).prepare_inputs()
    if o.alloc is None:
        o.allocate()
    res = {"step": step, "caches": int(count), "oracle_caches": int(o.count)}
    n = o.count
    res["alloc_exact"] = bool(count == n and np.array_equal(atlas, o.alloc["atlas"]) and
                              np.array_equal(entries[:n, :4].view(np.uint32), o.alloc["entries"][:n, :4].view(np.uint32)))
    # every step-th entry through the oracle's gather (a compact copy of their positions)
    idx = np.arange(0, n, step)
    sub = np.zeros((len(idx), entries.shape[1]), np.float32)
    sub[:, :4] = o.alloc["entries"][idx, :4]
    orc.light_caches(wl.constant, wl.volume, wl.spot_lights, o.vpls, o.blocks, o.chain, sub, 0, len(idx), wl.sh_order,
                     wl.indirect_shadow, False, threads)
    ok_sh, r_sh = close(entries[idx, 4:], sub[:, 4:])
    res.update(checked_entries=int(len(idx)), sh_ok=bool(ok_sh), sh_worst_ratio=float(r_sh),
               sh_max_abs=float(np.abs(sub[:, 4:]).max()) if len(idx) else 0.0)
    ok_img, r_img = True, 0.0
    if image is not None:
        e = np.ascontiguousarray(entries[:max(n, 1)], np.float32)
        img_o = orc.apply_caches(wl.constant, wl.per_frame, wl.volume, wl.transitions, wl.sh_order, wl.depth, wl.normal,
                                 wl.diffuse, o.alloc["atlas"], e, threads)
        a = np.asarray(image[..., :3], np.float64)
        b = np.asarray(img_o[..., :3], np.float64)
        tol = 1e-5 + 1e-3 * np.maximum(np.abs(a), np.abs(b))
        if image_is_half:
            tol = tol + np.maximum(np.abs(b) * 2.0 ** -11, 2.0 ** -25)  # one RGBA16F rounding (subnormal floor)
        err = np.abs(a - b)
        r_img = float(np.max(err / tol)) if err.size else 0.0
        ok_img = bool(np.all(err <= tol))
        res.update(checked_pixels=int(a.shape[0] * a.shape[1]), image_ok=ok_img, image_worst_ratio=r_img,
                   image_max=float(b.max()))
    res["worst_ratio"] = float(max(r_sh, r_img))
    res["ok"] = bool(res["alloc_exact"] and ok_sh and ok_img)
    return res
