#!/usr/bin/env python
"""Opcode histogram of one kernel from an ncu report's source page (needs --import-source on / -lineinfo).

    python tools/ncu_opcodes.py gpurun_out/frame.ncu-rep apply_kernel [units]   # units: warps to normalise by
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw))]
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    data = []
    for r in rows[hi + 1:]:
        if len(r) != len(h) or not r[h.index("Instructions Executed")].isdigit():
            break  # the next launch of the same kernel
        data.append(r)
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    ts = max(sum(int(r[isamp]) for r in data), 1)
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ia].split()
        o = toks[1] if toks[0].startswith("@") else toks[0]
        o = o.split(".")[0]
        op[o] += int(r[ie])
        ops[o] += int(r[isamp])
    print("# %s: %d warp instructions, %d SASS lines, %d samples" % (kern, tot, len(data), ts))
    for k, v in op.most_common(30):
        per = (" %8.2f /unit" % (v / units)) if units else ""
        print("%-10s %12d %5.1f%% inst %5.1f%% samples%s" % (k, v, 100.0 * v / tot, 100.0 * ops[k] / ts, per))


if __name__ == "__main__":
    main()
