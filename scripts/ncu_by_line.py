#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

usage: ncu_by_line.py report.ncu-rep cubin kernel_mangled_name [top]
Joins `ncu --page source --csv` (per-SASS-instruction counters) with
`nvdisasm -g` line annotations of the same cubin by instruction order.
"""
import csv, subprocess, sys, re, collections

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function section
start = next(i for i, l in enumerate(sass) if l.startswith("\t.section\t.text." + kname) or l.strip() == f".text.{kname}:")
lines = []  # (file,line) per instruction in order
cur = None
inl = None
for l in sass[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.sub(r"^_Z\d+|I.*$", "", kname)],
                     capture_output=True, text=True).stdout.splitlines()
# take the first matching kernel block
hdr_i = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
rdr = csv.reader(out[hdr_i:])
hdr = next(rdr)
ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0])
n = 0
tot = 0
for row in rdr:
    if not row or not row[0].startswith("0x"):
        break
    key = lines[n] if n < len(lines) else None
    n += 1
    ie = int(row[ci]); agg[key][0] += ie; agg[key][1] += int(row[si]); tot += ie
print(f"{n} SASS instructions, {len(lines)} disassembled, {tot} warp-instructions executed")
for k, (ie, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ie/tot:6.2f}%  {ie:12d}  samples {s:7d}  {k}")

# ---- attribute inlined helper code to the most recent kernel-body line (context) ----
if len(sys.argv) > 5:
    body_from = int(sys.argv[5])
    rdr = csv.reader(out[hdr_i:]); next(rdr)
    ctx = None; agg2 = collections.defaultdict(int); n = 0
    for row in rdr:
        if not row or not row[0].startswith("0x"):
            break
        key = lines[n] if n < len(lines) else None
        n += 1
        if key and key[0] == "cgb_sampler.cu" and key[1] >= body_from:
            ctx = key[1]
        agg2[ctx] += int(row[ci])
    print("---- by kernel-body context line ----")
    for k in sorted(agg2, key=lambda x: (x is None, x)):
        if agg2[k] * 200 > tot:
            print(f"{100*agg2[k]/tot:6.2f}%  line {k}")
