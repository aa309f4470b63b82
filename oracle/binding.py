"""ctypes binding of liboracle_drv.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. Inputs and outputs are numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from dynamicradiancevolume_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_drv.so")
_lib = None
_P = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build:
            raise ImportError("liboracle_drv.so missing: run `make -C oracle`")
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(LIB_PATH)
    lib.orc_default_threads.restype = C.c_int
    lib.orc_allocate_caches.restype = C.c_int
    lib.orc_allocated_cell_ids.restype = C.c_int
    lib.orc_cone_trace.restype = C.c_float
    lib.orc_sample_voxel.restype = C.c_float
    lib.orc_half_to_float.restype = C.c_float
    lib.orc_half_to_float.argtypes = [C.c_uint16]
    lib.orc_float_to_half.restype = C.c_uint16
    lib.orc_float_to_half.argtypes = [C.c_float]
    lib.orc_srgb8_to_linear.restype = C.c_float
    lib.orc_srgb8_to_linear.argtypes = [C.c_uint8]
    lib.orc_morton_decode_x.restype = C.c_uint32
    lib.orc_morton_decode_y.restype = C.c_uint32
    _lib = lib
    return lib


def default_threads():
    return load().orc_default_threads()


def allocate_caches(cb, pf, vi, transitions, depth, sh_order, max_caches, threads=0):
    """-> dict(count, atlas[z,y,x] u32, entries[max, stride/4] f32, counter, overflow, oob)."""
    lib = load()
    R, Cn = cb.AddressVolumeResolution, cb.NumAddressVolumeCascades
    atlas = np.zeros((R, R, R * Cn), dtype=np.uint32)
    stride = abi.entry_stride(sh_order)
    entries = np.zeros((max_caches, stride // 4), dtype=np.float32)
    counter = abi.CacheCounter()
    ov, oob = C.c_uint32(), C.c_uint32()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    n = lib.orc_allocate_caches(C.byref(cb), C.byref(pf), C.byref(vi), int(transitions), _ptr(depth), _ptr(atlas),
                                _ptr(entries), C.c_uint32(stride), C.c_uint32(max_caches), C.byref(counter),
                                C.byref(ov), C.byref(oob), int(threads))
    return dict(count=n, atlas=atlas, entries=entries, counter=counter, overflow=ov.value, oob=oob.value)


def allocated_cell_ids(cb, pf, vi, transitions, depth, threads=0):
    lib = load()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    n = lib.orc_allocated_cell_ids(C.byref(cb), C.byref(pf), C.byref(vi), int(transitions), _ptr(depth), None,
                                   C.c_uint32(0), int(threads))
    ids = np.zeros(n, dtype=np.int32)
    lib.orc_allocated_cell_ids(C.byref(cb), C.byref(pf), C.byref(vi), int(transitions), _ptr(depth), _ptr(ids),
                               C.c_uint32(n), int(threads))
    return ids


def rsm_downsample(flux, normal, depth):
    """One mip step: (flux[r,r,4] u16, normal[r,r,2] i16, depth[r,r,2] u16) -> half-resolution triple."""
    lib = load()
    r = flux.shape[0]
    h = r // 2
    fo = np.zeros((h, h, 4), np.uint16)
    no = np.zeros((h, h, 2), np.int16)
    do = np.zeros((h, h, 2), np.uint16)
    flux, normal, depth = (np.ascontiguousarray(a) for a in (flux, normal, depth))
    lib.orc_rsm_downsample(_ptr(flux), _ptr(normal), _ptr(depth), C.c_uint32(r), _ptr(fo), _ptr(no), _ptr(do))
    return fo, no, do


def rsm_mip_chain(flux, normal, depth):
    """Levels 0..log2(res)-1 as a list of triples (level 0 = the inputs)."""
    levels = [(flux, normal, depth)]
    while levels[-1][0].shape[0] > 2:
        levels.append(rsm_downsample(*levels[-1]))
    return levels


def generate_vpls(light, flux, normal, depth):
    lib = load()
    R = light.RSMReadResolution
    assert flux.shape[0] == R
    out = np.zeros(R * R, dtype=abi.VPL_DTYPE)
    flux, normal, depth = (np.ascontiguousarray(a) for a in (flux, normal, depth))
    lib.orc_generate_vpls(C.byref(light), _ptr(flux), _ptr(normal), _ptr(depth), _ptr(out))
    return out


def shadow_blocks(light, depth_lod):
    lib = load()
    R = light.RSMReadResolution
    n = R * R // light.IndirectShadowComputationSampleInterval
    out = np.zeros(n, dtype=abi.SHADOW_BLOCK_DTYPE)
    depth_lod = np.ascontiguousarray(depth_lod)
    assert depth_lod.shape[0] == R >> int(light.IndirectShadowComputationLod)
    lib.orc_shadow_blocks(C.byref(light), _ptr(depth_lod), _ptr(out))
    return out


def cone_trace(vi, chain, res, pos, block):
    lib = load()
    p = (C.c_float * 3)(*[float(x) for x in pos])
    b = np.ascontiguousarray(block)
    return lib.orc_cone_trace(C.byref(vi), _ptr(chain), C.c_uint32(res), p, _ptr(b))


def sample_voxel(chain, res, p, lod):
    lib = load()
    pp = (C.c_float * 3)(*[float(x) for x in p])
    return lib.orc_sample_voxel(_ptr(chain), C.c_uint32(res), pp, C.c_float(lod))


def light_caches(cb, vi, lights, vpls, blocks, voxel_chain, entries, first, count, sh_order, indirect_shadow,
                 fp64=False, threads=0):
    """In place on ``entries`` ([n, stride/4] f32). lights: list of abi.SpotLight; vpls/blocks: lists of arrays."""
    lib = load()
    n = len(lights)
    larr = (abi.SpotLight * n)(*lights)
    vp = (_P * n)(*[_ptr(v) for v in vpls])
    bp = (_P * n)(*[(_ptr(b) if b is not None else None) for b in (blocks or [None] * n)])
    assert entries.flags["C_CONTIGUOUS"] and entries.dtype == np.float32
    lib.orc_light_caches(C.byref(cb), C.byref(vi), larr, C.c_uint32(n), vp, bp, _ptr(voxel_chain), _ptr(entries),
                         C.c_uint32(entries.shape[1] * 4), C.c_uint32(first), C.c_uint32(count), int(sh_order),
                         int(bool(indirect_shadow)), int(bool(fp64)), int(threads))
    return entries


def apply_caches(cb, pf, vi, transitions, sh_order, depth, normal, diffuse, atlas, entries, threads=0):
    lib = load()
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.float32)
    depth, normal, diffuse, atlas = (np.ascontiguousarray(a) for a in (depth, normal, diffuse, atlas))
    lib.orc_apply_caches(C.byref(cb), C.byref(pf), C.byref(vi), int(transitions), int(sh_order), _ptr(depth),
                         _ptr(normal), _ptr(diffuse), _ptr(atlas), _ptr(entries), C.c_uint32(entries.shape[1] * 4),
                         C.c_uint32(entries.shape[0]), _ptr(out), int(threads))
    return out


def voxelize(vi, res, tris, world=None, target=None):
    lib = load()
    if target is None:
        target = np.zeros(res ** 3, np.uint8)
    w = np.eye(4, dtype=np.float32) if world is None else np.ascontiguousarray(world, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    lib.orc_voxelize(C.byref(vi), C.c_uint32(res), _ptr(tris), C.c_uint32(tris.size // 9), _ptr(w), _ptr(target))
    return target


def voxel_blend(volume, target, res, adaption):
    load().orc_voxel_blend(_ptr(volume), _ptr(target), C.c_uint32(res), C.c_float(adaption))
    return volume


def voxel_chain(level0, res):
    """Full mip chain (contiguous, level 0 first) from a level-0 volume."""
    chain = np.zeros(abi.voxel_chain_bytes(res), np.uint8)
    chain[: res ** 3] = level0
    load().orc_voxel_mips(_ptr(chain), C.c_uint32(res))
    return chain


def fill_rsm(light, position, normal, basecolor, coverage=None):
    """shader/fillrsm.frag:32-61 -> (flux[r,r,4] u16, normal[r,r,2] i16, depth[r,r,2] u16)."""
    lib = load()
    r = position.shape[0]
    fo = np.zeros((r, r, 4), np.uint16)
    no = np.zeros((r, r, 2), np.int16)
    do = np.zeros((r, r, 2), np.uint16)
    position, normal, basecolor = (np.ascontiguousarray(a, np.float32) for a in (position, normal, basecolor))
    cov = None if coverage is None else np.ascontiguousarray(coverage, np.uint8)
    lib.orc_fill_rsm(C.byref(light), _ptr(position), _ptr(normal), _ptr(basecolor), _ptr(cov), C.c_uint32(r), _ptr(fo),
                     _ptr(no), _ptr(do))
    return fo, no, do


def cone_trace_ao(pf, vi, chain, res, depth, normal, out=None, threads=0):
    """shader/ambientocclusion.frag:25-89 -> [H, W] float32 (``out`` keeps its values where the shader discards)."""
    lib = load()
    H, W = depth.shape
    if out is None:
        out = np.zeros((H, W), np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    normal = np.ascontiguousarray(normal, np.int16)
    lib.orc_cone_trace_ao(C.byref(pf), C.byref(vi), _ptr(chain), C.c_uint32(res), _ptr(depth), _ptr(normal), C.c_uint32(W),
                          C.c_uint32(H), _ptr(out), int(threads))
    return out


def tonemap(hdr_rgba, exposure, drago_divider):
    """shader/tonemapping.frag:21-31: [.., 4] float32 -> [.., 3] float32."""
    lib = load()
    hdr = np.ascontiguousarray(hdr_rgba, np.float32)
    n = hdr.size // 4
    out = np.zeros(hdr.shape[:-1] + (3,), np.float32)
    lib.orc_tonemap(_ptr(hdr), C.c_uint32(n), C.c_float(exposure), C.c_float(drago_divider), _ptr(out))
    return out


def half_to_float(h):
    return load().orc_half_to_float(int(h))


def float_to_half(f):
    return load().orc_float_to_half(float(f))


def pack_normal16i(n):
    nn = (C.c_float * 3)(*[float(x) for x in n])
    out = (C.c_int16 * 2)()
    load().orc_pack_normal16i(nn, out)
    return out[0], out[1]


def unpack_normal16i(px, py):
    i = (C.c_int16 * 2)(int(px), int(py))
    out = (C.c_float * 3)()
    load().orc_unpack_normal16i(i, out)
    return np.array(out[:], np.float32)


def srgb8_to_linear(v):
    return load().orc_srgb8_to_linear(int(v))


def morton_decode(k):
    lib = load()
    return lib.orc_morton_decode_x(C.c_uint32(k)), lib.orc_morton_decode_y(C.c_uint32(k))


# ---- f4: indirect specular (specular.cpp) ------------------------------------------------------------------
def specular_mip_texels(cb):
    """Texel count of the whole environment-map chain and the per-level sizes."""
    sizes, r, per = [], cb.SpecularEnvmapTotalSize, cb.SpecularEnvmapPerCacheSize_Texel
    while per >= 1:
        sizes.append(r)
        r //= 2
        per //= 2
    return sum(x * x for x in sizes), sizes


def light_caches_specular(cb, pf, vi, lights, vpls, blocks, voxel_chain, entries, count, sh_order, indirect_shadow):
    """In place on ``entries``; returns the mip chain buffer (uint32 R11G11B10F texels) with level 0 filled."""
    lib = load()
    n = len(lights)
    larr = (abi.SpotLight * n)(*lights)
    vp = (_P * n)(*[_ptr(v) for v in vpls])
    bp = (_P * n)(*[(_ptr(b) if b is not None else None) for b in (blocks or [None] * n)])
    total, _ = specular_mip_texels(cb)
    mips = np.zeros(total, np.uint32)
    lib.orc_light_caches_specular(C.byref(cb), C.byref(pf), C.byref(vi), larr, C.c_uint32(n), vp, bp, _ptr(voxel_chain),
                                  _ptr(entries), C.c_uint32(entries.shape[1] * 4), C.c_uint32(count), int(sh_order),
                                  int(bool(indirect_shadow)), _ptr(mips))
    return mips


def specular_mips(cb, count, mips):
    load().orc_specular_mips(C.byref(cb), C.c_uint32(count), _ptr(mips))
    return mips


def specular_fill_holes(cb, count, max_level, mips):
    load().orc_specular_fill_holes(C.byref(cb), C.c_uint32(count), C.c_uint32(max_level), _ptr(mips))
    return mips


def apply_caches_specular(cb, pf, vi, transitions, sh_order, depth, normal, diffuse, rough_metal, atlas, entries, mips,
                          threads=0):
    lib = load()
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.float32)
    depth, normal, diffuse, rough_metal, atlas = (np.ascontiguousarray(a) for a in (depth, normal, diffuse, rough_metal, atlas))
    lib.orc_apply_caches_specular(C.byref(cb), C.byref(pf), C.byref(vi), int(bool(transitions)), int(sh_order), _ptr(depth),
                                  _ptr(normal), _ptr(diffuse), _ptr(rough_metal), _ptr(atlas), _ptr(entries),
                                  C.c_uint32(entries.shape[1] * 4), C.c_uint32(entries.shape[0]), _ptr(mips), _ptr(out),
                                  int(threads))
    return out
