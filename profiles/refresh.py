#!/usr/bin/env python
"""Regenerate the text artifacts of profiles/ from one ncu report of the bench step.

usage: refresh.py <report.ncu-rep> <launches.csv> <tag>

Needs ncu, cuobjdump, nvdisasm and the in-tree .so that was profiled (-lineinfo build).
Writes   <tag>_step_kernels.txt       counter summary of every kernel in the report (summarize.py)
         <tag>_<kernel>_lines.txt     per-source-line instruction counts / stall samples (line_profile.py)
         <tag>_launch_shares.txt      kernel shares of one step from the launch list
         <tag>_sass_<kernel>.sass.gz  cuobjdump -sass of every kernel, <tag>_sass_summary.txt mnemonic histograms
         ncu_traffic.json             DRAM bytes per launch, read by bench.py for roofline.traffic
"""
import collections
import csv
import gzip
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(ROOT, "ribotricer_b200", "libribotricer_b200.so")
rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
py = sys.executable

# 1. counter summary
txt = subprocess.run([py, os.path.join(HERE, "summarize.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(HERE, f"{tag}_step_kernels.txt"), "w").write(txt)

# 2. per-line profiles: (mangled symbol substring, ncu regex, file tag)
KERNELS = [("bin_stream_kernelILb1ELb1E", "bin_stream", "bin_stream_kernel"),
           ("atom_pass_kernelILi2ELb0E", "atom_pass", "atom_pass_kernel"),
           ("compose_refs_kernelILb0E", "compose_refs", "compose_refs_kernel")]
for sym, rx, name in KERNELS:
    out = subprocess.run([py, os.path.join(HERE, "line_profile.py"), rep, sym, "45", rx], capture_output=True, text=True)
    open(os.path.join(HERE, f"{tag}_{name}_lines.txt"), "w").write(out.stdout + out.stderr)

# 3. DRAM traffic per launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = re.sub(r"^void |<.*|\(.*", "", name)

    def val(metric):
        i = hdr.index(metric)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return int(float(r[i].replace(",", "")) * scale)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    traffic[key] = {"workload": "C2", "layout": "compact", "dram_bytes": rd + wr, "dram_read_bytes": rd,
                    "dram_write_bytes": wr,
                    "source": f"profiles/{tag}_step_kernels.txt (ncu --set full, launch of `{name}` inside bench.py's step)"}
old = {}
try:
    old = json.load(open(os.path.join(HERE, "ncu_traffic.json")))
except Exception:
    pass
old.update(traffic)
json.dump(old, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)

# 4. kernel shares of one step
per = collections.OrderedDict()
n_steps = 0
for r in csv.reader(open(launches)):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        k = re.sub(r"^void |rt::|<.*|\(.*", "", r[4])
        grid = int(r[8].strip("()").split(",")[0])
        per.setdefault(k, []).append((grid, float(r[14]) / 1e6))
STEP = ("zone_bounds_kernel", "zone_carry_kernel", "bin_stream_kernel", "zone_spill_kernel", "bin_psites_kernel", "atom_pass_kernel",
        "compose_refs_kernel", "score_orfs_kernel")
per = collections.OrderedDict((k, v) for k, v in per.items() if k in STEP)   # not the one-time set-up kernels
for k in per:   # the e2e leg launches K1 in 1 M-read chunks: keep the whole-library launches only
    g = max(x[0] for x in per[k])
    per[k] = [t for x, t in per[k] if x == g]
    # persistent kernels keep their grid whatever the work: the e2e leg scores in four parts (rt_score_host), drop those
    longest = max(per[k])
    per[k] = [t for t in per[k] if t >= 0.6 * longest]
lines = []
tot = sum(sum(v) / max(1, len(v)) for v in per.values())
for k, v in per.items():
    m = sum(v) / max(1, len(v))
    lines.append(f"{k:28s} {m:8.4f} ms  {100 * m / tot:5.1f} %   ({len(v)} launches averaged, serialised, cold cache)")
lines.append(f"{'sum':28s} {tot:8.4f} ms  (the step of bench.py: rt_bin_stream_fresh + rt_score; no clear of the buffer)")
open(os.path.join(HERE, f"{tag}_launch_shares.txt"), "w").write("\n".join(lines) + "\n")

# 5. SASS
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
summary = []
for f in funcs:
    sym = f.split("\n", 1)[0].strip()
    m = re.search(r"_ZN2rt\d+([a-z_0-9]+?_kernel)(I[A-Za-z0-9]+?E)?E", sym)
    name = (m.group(1) + ("_" + m.group(2)[1:-1] if m.group(2) else "")) if m else sym
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f, re.M)
    hist = collections.Counter(ops)
    summary.append(f"{name}: {len(ops)} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
    with gzip.open(os.path.join(HERE, f"{tag}_sass_{name}.sass.gz"), "wt") as fh:
        fh.write("Function : " + f)
open(os.path.join(HERE, f"{tag}_sass_summary.txt"), "w").write("\n".join(summary) + "\n")
print("\n".join(lines))
