"""bench.py's reference arm (the CPU oracle leg, the only leg that runs without a GPU) prints exactly one JSON line
with the keys the driver reads; the b200 arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--config", "0", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GI ms/frame @1080p,16k VPLs" and d["unit"] == "ms/frame"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--config", "0", "--steps", "1", "--warmup", "0", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_never_maps_the_cuda_library():
    """The CPU arm builds its workload through libdrv_host.so (the packers compiled with g++): libdrv_gi.so, the
    product, must not be among the objects the process has mapped after a whole reference frame."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import workloads\n"
            "from oracle.frame import OracleFrame\n"
            "wl = workloads.config(0, width=64, height=64).build()\n"
            "o = OracleFrame(wl, threads=2); o.prepare_inputs(); o.frame()\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libdrv_host.so' in maps and 'liboracle_drv.so' in maps, 'helper libraries not mapped'\n"
            "assert 'libdrv_gi.so' not in maps, 'the CUDA library was loaded by the CPU arm'\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
