#!/usr/bin/env python
"""Print every drv_microbench figure (roofline denominators + the operand-delivery study) as one JSON object."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynamicradiancevolume_b200 as drv

lib = drv.load()
out = {}
for w in range(lib.drv_microbench_count()):
    r = C.c_double()
    if lib.drv_microbench(0, w, C.byref(r)) == 0:
        out[lib.drv_microbench_name(w).decode()] = r.value
print(json.dumps(out, indent=1))
