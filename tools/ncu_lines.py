#!/usr/bin/env python
"""Per-source-line dynamic instruction counts of one kernel: joins the per-SASS `Instructions Executed` of an ncu
report (source page) with the line table of the object file the report was captured from (nvdisasm -g).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep cone_kernel dynamicradiancevolume_b200/csrc/gather.o [units]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_counts(rep, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    ie, isamp, ia = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    out = []
    for r in rows[hi + 1:]:
        if len(r) != len(h) or not r[ie].isdigit():
            break
        out.append((r[ia], int(r[ie]), int(r[isamp])))
    return out


def line_table(obj, kern):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for l in txt.split("\n"):
        if l.startswith("//---") and ".text." in l:
            on = kern in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, kern, obj = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else None
    sc = sass_counts(rep, kern)
    lt = line_table(obj, kern)
    if len(sc) != len(lt):
        sys.stderr.write("warning: %d SASS lines in the report, %d in the object (different build?)\n" % (len(sc), len(lt)))
    agg, smp = collections.Counter(), collections.Counter()
    for (src, n, s), ln in zip(sc, lt):
        agg[ln] += n
        smp[ln] += s
    tot, ts = sum(agg.values()), max(sum(smp.values()), 1)
    cache = {}
    print("# %s: %d warp instructions" % (kern, tot))
    for ln, n in agg.most_common(40):
        text = ""
        if ln:
            path = os.path.join(os.path.dirname(os.path.abspath(obj)), ln[0])
            if path not in cache and os.path.exists(path):
                cache[path] = open(path).read().split("\n")
            if path in cache and ln[1] <= len(cache[path]):
                text = cache[path][ln[1] - 1].strip()[:110]
        per = (" %7.2f/unit" % (n / units)) if units else ""
        print("%5.1f%% inst %5.1f%% smp%s  %s:%s  %s" % (100.0 * n / tot, 100.0 * smp[ln] / ts, per, ln[0] if ln else "?", ln[1] if ln else "?", text))


if __name__ == "__main__":
    main()
