"""Builds libdrv_gi.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libdrv_gi.so")
HOST_LIB = os.path.join(HERE, "libdrv_host.so")  # the drv_pack_* entry points alone, plain g++ (see _lib.load_host)
SOURCES = ["ctx.cu", "alloc.cu", "rsm.cu", "voxel.cu", "gather.cu", "apply.cu", "adjacent.cu", "specular.cu", "microbench.cu", "host_pack.cpp"]
HEADERS = ["ctx.h", "device_math.cuh", "voxel_sample.cuh", os.path.join(ROOT, "include", "drv_gi.h"), os.path.join(ROOT, "include", "drv_math.h"), os.path.join(ROOT, "include", "drv_r11g11b10.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _compile(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
    deps = [path] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    if _mtime(obj) > max(_mtime(d) for d in deps):
        return obj
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a and link libdrv_gi.so. Returns the library path."""
    if force:
        for s in SOURCES:
            o = os.path.join(CSRC, os.path.splitext(s)[0] + ".o")
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if _mtime(LIB) < max(_mtime(o) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_host()
    return LIB


def build_host():
    src = os.path.join(CSRC, "host_pack.cpp")
    deps = [src, os.path.join(ROOT, "include", "drv_gi.h"), os.path.join(ROOT, "include", "drv_math.h")]
    if _mtime(HOST_LIB) > max(_mtime(d) for d in deps):
        return HOST_LIB
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", "-o", HOST_LIB, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for libdrv_host.so:\n%s\n%s" % (r.stdout, r.stderr))
    return HOST_LIB


def build_aux():
    """Build the test/bench helper libraries (oracle, scenes) and the C++ test of the host-side mirror
    (tests/cpp: drv::Renderer of include/drv_renderer.hpp against the oracle) with make."""
    for d in ("oracle", "scenes", os.path.join("tests", "cpp")):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, d)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("make -C %s failed:\n%s\n%s" % (d, r.stdout, r.stderr))


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    build_aux()
