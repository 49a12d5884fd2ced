#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info -> per-source-line instruction counts.

usage: line_profile.py <report.ncu-rep> <kernel-symbol-substring> [top] [ncu-kernel-regex]
(the regex selects the kernel inside a report that holds several, e.g. "bin_psites")
Needs: ncu, cuobjdump, nvdisasm; the .so must be the one that was profiled (-lineinfo build).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, sym = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("LINE_PROFILE_SO", os.path.join(root, "ribotricer_b200", "libribotricer_b200.so"))
with tempfile.TemporaryDirectory() as tmp:
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# per-function list of (line, sass text)
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = sym in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2).strip()))
flt = ["-k", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
H = rows[hdr]
ie, smp = H.index("Instructions Executed"), H.index("# Samples")
sass = rows[hdr + 1:]
end = next((i for i, r in enumerate(sass) if r and r[0] == 'Kernel Name'), len(sass))   # a report may repeat the kernel
sass = sass[:end]
assert len(sass) == len(lines), (len(sass), len(lines))
src = open(os.environ.get("LINE_PROFILE_SRC", os.path.join(root, "ribotricer_b200", "csrc", "rt_kernels.cuh"))).read().splitlines()
agg = {}
tot_i = tot_s = 0
for (line, _), r in zip(lines, sass):
    n, s = int(r[ie] or 0), int(r[smp] or 0)
    a = agg.setdefault(line, [0, 0])
    a[0] += n
    a[1] += s
    tot_i += n
    tot_s += s
print(f"total warp instructions {tot_i}, samples {tot_s}")
for line, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    fname, ln = line if line else ("?", 0)
    text = src[ln - 1].strip() if fname == "rt_kernels.cuh" and 0 < ln <= len(src) else fname
    print(f"{n:>13} {100 * n / tot_i:5.1f}%  smp {100 * s / max(1, tot_s):5.1f}%  L{ln}: {text[:100]}")
