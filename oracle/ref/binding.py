"""ctypes binding of oracle/_ref/libdrv_ref.so — the reference's own GLSL programs, rewritten mechanically into C++
(oracle/ref/glsl2cpp.py) and compiled with g++ (`make -C oracle ref`). TEST INFRASTRUCTURE ONLY: it exists to pin
oracle/ to the reference's source text (tests/test_oracle_vs_ref.py); nothing in the product path may import it.

The library is built only where /root/reference exists (this container); the GPU box gets the prebuilt file.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from dynamicradiancevolume_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.dirname(_HERE)
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libdrv_ref.so")
REFERENCE_SHADERS = "/root/reference/DynamicRadianceVolume/shader"
_lib = None
_P = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def available():
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE_SHADERS)


def build():
    """`make -C oracle ref` — needs the reference checkout; a no-op when the library is up to date."""
    if not os.path.isdir(REFERENCE_SHADERS):
        return os.path.exists(LIB_PATH)
    r = subprocess.run(["make", "-C", ORACLE_DIR, "-j8", "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("make -C oracle ref failed:\n%s\n%s" % (r.stdout[-4000:], r.stderr[-4000:]))
    return True


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    if not os.path.exists(LIB_PATH):
        raise ImportError("oracle/_ref/libdrv_ref.so is missing and /root/reference is not available to build it")
    lib = C.CDLL(LIB_PATH)
    lib.ref_allocate_caches.restype = C.c_int
    lib.ref_source.restype = C.c_char_p
    _lib = lib
    return lib


def allocate_caches(cb, pf, vi, transitions, depth, sh_order, threads=0):
    """shader/cacheGather.comp + cachePrepareLighting.comp. entries holds one slot per CAV cell (the shader has no
    capacity check). -> dict(count, atlas[z,y,x], entries[cells, stride/4], counter)."""
    lib = load()
    R, Cn = cb.AddressVolumeResolution, cb.NumAddressVolumeCascades
    cells = R * R * R * Cn
    atlas = np.zeros((R, R, R * Cn), np.uint32)
    stride = abi.entry_stride(sh_order)
    entries = np.zeros((cells, stride // 4), np.float32)
    counter = abi.CacheCounter()
    depth = np.ascontiguousarray(depth, np.float32)
    n = lib.ref_allocate_caches(C.byref(cb), C.byref(pf), C.byref(vi), int(bool(transitions)), int(sh_order), _ptr(depth),
                                _ptr(atlas), _ptr(entries), C.c_uint32(cells), C.byref(counter), int(threads))
    return dict(count=n, atlas=atlas, entries=entries, counter=counter)


def light_caches(cb, pf, vi, lights, rsm_levels, read_levels, voxel_chain, voxel_res, entries, count, sh_order,
                 indirect_shadow, tap=False, threads=0):
    """shader/cacheLightingRSM.comp, one dispatch per light, in place on ``entries``.
    rsm_levels[l] = list of (flux, normal, depthLinSq) per mip level; read_levels[l] = rsmReadLod of light l.
    tap=True also returns, per light, the VPL list the shader derived in shared memory ([R^2, 10] float32:
    Flux 3, DiscArea, Position 3, Normal 3)."""
    lib = load()
    n = len(lights)
    larr = (abi.SpotLight * n)(*lights)
    keep = []
    flux_p, normal_p, depth_pp, nlev = (_P * n)(), (_P * n)(), (_P * n)(), (C.c_uint32 * n)()
    taps = []
    tap_p = (_P * n)()
    for i in range(n):
        lv = rsm_levels[i]
        rl = read_levels[i]
        f, nm = np.ascontiguousarray(lv[rl][0]), np.ascontiguousarray(lv[rl][1])
        ds = [np.ascontiguousarray(l[2]) for l in lv[rl:]]
        dptr = (_P * len(ds))(*[_ptr(d) for d in ds])
        keep += [f, nm, ds, dptr]
        flux_p[i], normal_p[i] = _ptr(f), _ptr(nm)
        depth_pp[i] = C.cast(dptr, _P)
        nlev[i] = len(ds)
        R = int(lights[i].RSMReadResolution)
        t = np.zeros((R * R, 10), np.float32)
        taps.append(t)
        tap_p[i] = _ptr(t)
    assert entries.flags["C_CONTIGUOUS"] and entries.dtype == np.float32
    assert entries.shape[1] * 4 == abi.entry_stride(sh_order)
    lib.ref_light_caches(C.byref(cb), C.byref(pf), C.byref(vi), larr, C.c_uint32(n), flux_p, normal_p, depth_pp, nlev,
                         _ptr(voxel_chain), C.c_uint32(voxel_res), _ptr(entries), C.c_uint32(count), int(sh_order),
                         int(bool(indirect_shadow)), tap_p if tap else None, int(threads))
    return taps if tap else entries


def apply_caches(cb, pf, vi, transitions, sh_order, depth, normal, diffuse, atlas, entries, threads=0):
    lib = load()
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.float32)
    depth, normal, diffuse, atlas = (np.ascontiguousarray(a) for a in (depth, normal, diffuse, atlas))
    lib.ref_apply_caches(C.byref(cb), C.byref(pf), C.byref(vi), int(bool(transitions)), int(sh_order), _ptr(depth), _ptr(normal),
                         _ptr(diffuse), _ptr(atlas), _ptr(entries), C.c_uint32(entries.shape[0]), _ptr(out), int(threads))
    return out


def voxel_blend(volume, target, res, adaption):
    load().ref_voxel_blend(_ptr(volume), _ptr(target), C.c_uint32(res), C.c_float(adaption))
    return volume


def voxel_chain(level0, res):
    chain = np.zeros(abi.voxel_chain_bytes(res), np.uint8)
    chain[: res ** 3] = level0
    load().ref_voxel_mips(_ptr(chain), C.c_uint32(res))
    return chain


def rsm_downsample(flux, normal, depth):
    lib = load()
    r = flux.shape[0]
    h = r // 2
    fo, no, do = np.zeros((h, h, 4), np.uint16), np.zeros((h, h, 2), np.int16), np.zeros((h, h, 2), np.uint16)
    flux, normal, depth = (np.ascontiguousarray(a) for a in (flux, normal, depth))
    lib.ref_rsm_downsample(_ptr(flux), _ptr(normal), _ptr(depth), C.c_uint32(r), _ptr(fo), _ptr(no), _ptr(do))
    return fo, no, do


def fill_rsm(light, position, normal, basecolor, coverage=None):
    """shader/fillrsm.frag -> (flux[r,r,4] u16, normal[r,r,2] i16, depth[r,r,2] u16)."""
    lib = load()
    r = position.shape[0]
    fo, no, do = np.zeros((r, r, 4), np.uint16), np.zeros((r, r, 2), np.int16), np.zeros((r, r, 2), np.uint16)
    position, normal, basecolor = (np.ascontiguousarray(a, np.float32) for a in (position, normal, basecolor))
    cov = None if coverage is None else np.ascontiguousarray(coverage, np.uint8)
    lib.ref_fill_rsm(C.byref(light), _ptr(position), _ptr(normal), _ptr(basecolor), _ptr(cov), C.c_uint32(r), _ptr(fo),
                     _ptr(no), _ptr(do))
    return fo, no, do


def cone_trace_ao(cb, pf, vi, chain, res, depth, normal, out=None, threads=0):
    """shader/ambientocclusion.frag -> [H, W] float32 (``out`` keeps its values where the shader discards)."""
    lib = load()
    H, W = depth.shape
    if out is None:
        out = np.zeros((H, W), np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    normal = np.ascontiguousarray(normal, np.int16)
    lib.ref_cone_trace_ao(C.byref(cb), C.byref(pf), C.byref(vi), _ptr(chain), C.c_uint32(res), _ptr(depth), _ptr(normal),
                          C.c_uint32(W), C.c_uint32(H), _ptr(out), int(threads))
    return out


def tonemap(hdr_rgba, exposure, drago_divider):
    lib = load()
    hdr = np.ascontiguousarray(hdr_rgba, np.float32)
    n = hdr.size // 4
    out = np.zeros(hdr.shape[:-1] + (3,), np.float32)
    lib.ref_tonemap(_ptr(hdr), C.c_uint32(n), C.c_float(exposure), C.c_float(drago_divider), _ptr(out))
    return out


# ---- f4: INDIRECT_SPECULAR variants ----------------------------------------------------------------------------
def light_caches_specular(cb, pf, vi, lights, rsm_levels, read_levels, voxel_chain, voxel_res, entries, count, sh_order,
                          indirect_shadow, total_texels):
    """cacheLightingRSM.comp with INDIRECT_SPECULAR + DIRECT_SPECULAR_MAP_WRITE (per-cache size 16), work groups in
    order on one thread. Built variants: (SH1, unshadowed) and (SH2, shadowed). Returns the mip buffer, level 0 filled."""
    lib = load()
    n = len(lights)
    larr = (abi.SpotLight * n)(*lights)
    keep = []
    flux_p, normal_p, depth_pp, nlev = (_P * n)(), (_P * n)(), (_P * n)(), (C.c_uint32 * n)()
    for i in range(n):
        lv, rl = rsm_levels[i], read_levels[i]
        f, nm = np.ascontiguousarray(lv[rl][0]), np.ascontiguousarray(lv[rl][1])
        ds = [np.ascontiguousarray(l[2]) for l in lv[rl:]]
        dptr = (_P * len(ds))(*[_ptr(d) for d in ds])
        keep += [f, nm, ds, dptr]
        flux_p[i], normal_p[i] = _ptr(f), _ptr(nm)
        depth_pp[i] = C.cast(dptr, _P)
        nlev[i] = len(ds)
    mips = np.zeros(total_texels, np.uint32)
    lib.ref_light_caches_specular.restype = C.c_int
    rc = lib.ref_light_caches_specular(C.byref(cb), C.byref(pf), C.byref(vi), larr, C.c_uint32(n), flux_p, normal_p, depth_pp,
                                       nlev, _ptr(voxel_chain), C.c_uint32(voxel_res), _ptr(entries), C.c_uint32(count),
                                       int(sh_order), int(bool(indirect_shadow)), _ptr(mips))
    if rc != 0:
        raise ValueError("this INDIRECT_SPECULAR variant is not built into oracle/_ref")
    return mips


def specular_mips(cb, count, mips):
    load().ref_specular_mips(C.byref(cb), C.c_uint32(count), _ptr(mips))
    return mips


def specular_fill_holes(cb, count, max_level, mips):
    load().ref_specular_fill_holes(C.byref(cb), C.c_uint32(count), C.c_uint32(max_level), _ptr(mips))
    return mips


def apply_caches_specular(cb, pf, vi, transitions, sh_order, depth, normal, diffuse, rough_metal, atlas, entries, mips,
                          threads=0):
    lib = load()
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.float32)
    depth, normal, diffuse, rough_metal, atlas = (np.ascontiguousarray(a) for a in (depth, normal, diffuse, rough_metal, atlas))
    lib.ref_apply_caches_specular.restype = C.c_int
    rc = lib.ref_apply_caches_specular(C.byref(cb), C.byref(pf), C.byref(vi), int(bool(transitions)), int(sh_order), _ptr(depth),
                                       _ptr(normal), _ptr(diffuse), _ptr(rough_metal), _ptr(atlas), _ptr(entries),
                                       C.c_uint32(entries.shape[0]), _ptr(mips), _ptr(out), int(threads))
    if rc != 0:
        raise ValueError("this INDIRECT_SPECULAR variant is not built into oracle/_ref")
    return out
