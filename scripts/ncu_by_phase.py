#!/usr/bin/env python
"""Where the warps of cgb_sample_kernel spend their time: executed warp-instructions and warp-state
samples of an `ncu --set full --import-source on` capture, aggregated by the phase of the kernel
(phase 0 / trial loop / cooperative finishes / phase 2 / final flush) and, per phase, by stall reason.

usage: ncu_by_phase.py report.ncu-rep [out.txt]

Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -g` line
annotations of the cubin extracted from the built library, by instruction order.  Code inlined
from helpers (Philox, the screen, sincos) is attributed to the kernel-body line that precedes it.
The phase boundaries are found in cgb_sampler.cu by their marker comments / lambdas.
"""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "cluster_generator_b200", "csrc", "cgb_sampler.cu")
SO = os.path.join(ROOT, "cluster_generator_b200", "lib", "libcgb200.so")
KERNEL = "_Z17cgb_sample_kernelIdLi0ELb0EEv12SampleParams"
rep = sys.argv[1]
out_path = sys.argv[2] if len(sys.argv) > 2 else None

src = open(SRC).read().splitlines()
def line_of(needle, start=0):
    return next(i + 1 for i, l in enumerate(src) if i >= start and needle in l)
k0 = line_of("cgb_sample_kernel(const __grid_constant__ SampleParams p)")
marks = [("phase 0 (radius, psi, envelope)", line_of("auto acquire = [&]", k0)),
         ("phase 2 (directions, stores, moments)", line_of("auto epilogue = [&]", k0)),
         ("cooperative finishes", line_of("auto coop_finish = [&]", k0)),
         ("phase 1 lean loop (Philox + screens)", line_of("auto finish_lanes = [&]", k0)),
         ("final flush + end-of-kernel barrier", line_of("// ---------------- moments ----------------", k0)),
         (None, line_of("// Grid-wide pass over the deferred particles", k0))]
ranges = [(marks[i][0], marks[i][1], marks[i + 1][1] - 1) for i in range(len(marks) - 1)]

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    cubin = next(os.path.join(tmp, f) for f in os.listdir(tmp) if "cgb_sampler" in f)
    sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith("\t.section\t.text." + KERNEL) or l.strip() == f".text.{KERNEL}:")
lines, cur = [], None
for l in sass[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)

page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:cgb_sample_kernel"],
                      capture_output=True, text=True).stdout.splitlines()
hdr_i = next(i for i, l in enumerate(page) if l.startswith('"Address"'))
rdr = csv.reader(page[hdr_i:])
hdr = next(rdr)
ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
rows = []
for row in rdr:
    if not row or not row[0].startswith("0x"):
        break
    rows.append(row)
rows = rows[:len(lines)]  # the first captured launch
assert len(rows) == len(lines), f"{len(rows)} SASS rows in the report, {len(lines)} in the library: capture and build differ"

agg = collections.OrderedDict((name, [0, 0, collections.Counter()]) for name, _, _ in ranges)
agg["other"] = [0, 0, collections.Counter()]
ctx = None
for key, row in zip(lines, rows):
    if key and key[0] == "cgb_sampler.cu" and key[1] >= ranges[0][1]:
        ctx = key[1]
    name = next((nm for nm, a, b in ranges if ctx and a <= ctx <= b), "other")
    agg[name][0] += int(row[ci]); agg[name][1] += int(row[si])
    for i in stall_cols:
        if row[i] and row[i] != "0":
            agg[name][2][hdr[i][6:]] += int(row[i])
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
out = [f"cgb_sample_kernel<double,0,false>: {len(rows)} SASS instructions, {ti} executed warp-instructions, {ts} warp-state samples",
       f"(report {os.path.basename(rep)}; phases located in cgb_sampler.cu lines {ranges[0][1]}-{ranges[-1][2]})", "",
       f"{'region':42s} {'instructions':>12s} {'samples':>8s}  stall reasons (share of the region's samples)"]
for name, (i, s, c) in agg.items():
    t = max(sum(c.values()), 1)
    out.append(f"{name:42s} {100 * i / ti:11.1f}% {100 * s / ts:7.1f}%  " +
               ", ".join(f"{k} {100 * v / t:.0f}%" for k, v in c.most_common(5)))
text = "\n".join(out) + "\n"
print(text, end="")
if out_path:
    open(out_path, "w").write(text)
