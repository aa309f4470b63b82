"""ctypes binding of libdrv_scenes.so: procedural synthetic inputs (see scenes/scenes.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

from dynamicradiancevolume_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrv_scenes.so")
_lib = None
_P = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_P)


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build:
            raise ImportError("libdrv_scenes.so missing: run `make -C scenes`")
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(LIB_PATH)
    lib.scn_create.restype = _P
    lib.scn_create.argtypes = [C.c_char_p, C.c_float]
    lib.scn_destroy.argtypes = [_P]
    lib.scn_num_boxes.restype = C.c_uint32
    lib.scn_num_boxes.argtypes = [_P]
    lib.scn_bounding_box.argtypes = [_P, _P, _P]
    lib.scn_triangles.restype = C.c_uint32
    lib.scn_triangles.argtypes = [_P, _P, C.c_uint32]
    lib.scn_render_gbuffer.argtypes = [_P, _P, C.c_uint32, C.c_uint32, _P, _P, _P, C.c_int]
    lib.scn_render_rsm.argtypes = [_P, _P, _P, _P, _P, C.c_int]
    lib.scn_sweep_entries.argtypes = [C.c_uint32, C.c_uint32, _P]
    lib.scn_sweep_vpls.argtypes = [C.c_uint32, C.c_uint32, C.c_float, _P]
    _lib = lib
    return lib


class SceneGeometry:
    """A procedural box scene ("cornell" or "atrium")."""

    def __init__(self, name: str, scale: float = 1.0):
        self.lib = load()
        self.name = name
        self.handle = self.lib.scn_create(name.encode(), scale)
        if not self.handle:
            raise ValueError("unknown scene %r" % name)

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.scn_destroy(self.handle)
            self.handle = None

    def bounding_box(self):
        mn = np.zeros(3, np.float32)
        mx = np.zeros(3, np.float32)
        self.lib.scn_bounding_box(self.handle, _ptr(mn), _ptr(mx))
        return mn, mx

    def triangles(self) -> np.ndarray:
        n = self.lib.scn_triangles(self.handle, None, 0)
        out = np.zeros((n, 9), np.float32)
        self.lib.scn_triangles(self.handle, _ptr(out), n)
        return out

    def render_gbuffer(self, per_frame: abi.PerFrame, width: int, height: int, threads: int = 0):
        depth = np.zeros((height, width), np.float32)
        normal = np.zeros((height, width, 2), np.int16)
        diffuse = np.zeros((height, width, 4), np.uint8)
        self.lib.scn_render_gbuffer(self.handle, C.addressof(per_frame), width, height, _ptr(depth), _ptr(normal),
                                    _ptr(diffuse), threads)
        return depth, normal, diffuse

    def render_rsm(self, light: abi.SpotLight, threads: int = 0):
        r = light.RSMRenderResolution
        flux = np.zeros((r, r, 4), np.uint16)
        normal = np.zeros((r, r, 2), np.int16)
        depth = np.zeros((r, r, 2), np.uint16)
        self.lib.scn_render_rsm(self.handle, C.addressof(light), _ptr(flux), _ptr(normal), _ptr(depth), threads)
        return flux, normal, depth


def sweep_entries(seed: int, n: int) -> np.ndarray:
    out = np.zeros((n, 4), np.float32)
    load().scn_sweep_entries(seed & 0xFFFFFFFF, n, _ptr(out))
    return out


def sweep_vpls(seed: int, n: int, val_area_factor: float) -> np.ndarray:
    out = np.zeros(n, dtype=abi.VPL_DTYPE)
    load().scn_sweep_vpls(seed & 0xFFFFFFFF, n, val_area_factor, _ptr(out))
    return out
