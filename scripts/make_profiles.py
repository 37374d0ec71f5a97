#!/usr/bin/env python
"""Summarise gpurun_out/{launches_rNN.csv, prof_rNN.ncu-rep, bench_rNN.json} into profiles/.
usage: make_profiles.py r01"""
import collections, csv, json, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = "gpurun_out"
lines = [l for l in open(f"{G}/launches_{tag}.csv") if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    a = agg.setdefault(row["Kernel Name"], [0, 0.0])
    a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list, {tag} (`ncu --metrics gpu__time_duration.sum --clock-control none -c 500 python bench.py --steps 30 --warmup 3`)", "",
       "Per-launch times under ncu are cold-cache and serialised: compare SHARES. First 500 launches = table build + f(E) builds + warm-up + 30 timed steps (+ part of the e2e leg).", "",
       "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:90]}` | {a[0]} | {a[1]/1e3:.1f} | {a[1]/a[0]/1e3:.2f} | {100*a[1]/tot:.1f}% |")
st = {k: a for k, a in agg.items() if any(x in k for x in ("cgb_sample_kernel", "cgb_gas_kernel", "cgb_hard_kernel"))}
ts = sum(a[1] / a[0] for a in st.values())
out += ["", "## one bench step (DM sample kernel + deferred pass + gas kernel)", "", "| kernel | avg us | share of step |", "|---|---:|---:|"]
for k, a in st.items():
    out.append(f"| `{k[:70]}` | {a[1]/a[0]/1e3:.1f} | {100*(a[1]/a[0])/ts:.1f}% |")
b = json.load(open(f"{G}/bench_{tag}.json"))
out += ["", f"bench.py (not under ncu, CUDA events): ms_per_step {b['ms_per_step']:.3f}, sample kernel "
        f"{b['roofline']['kernel_ms']:.3f} ms = {100*b['roofline']['kernel_ms']/b['ms_per_step']:.1f}% of the step -- agrees with the share above."]
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")
shutil.copy(f"{G}/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
shutil.copy(f"{G}/bench_{tag}.json", f"profiles/{tag}_bench.json")

raw = subprocess.run(["ncu", "-i", f"{G}/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, units = r[0], r[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__warps_eligible.avg.per_cycle_active']
o2 = [f"# ncu --set full, {tag} (`-k regex:cgb_sample_kernel|cgb_gas_kernel|cgb_eddington_g_kernel -c 10`, bench.py --steps 3 --warmup 3; last capture of each kernel)", "",
      "Raw report kept out of git (9 MB); regenerate with the command above. Values per launch, workload = bench.py (1e7 Vikhlinin DM / 1e7 gas).", ""]
last = collections.OrderedDict((row[h.index('Kernel Name')], row) for row in r[2:])  # one (the last) capture per kernel
for row in last.values():
    o2 += [f"## {row[h.index('Kernel Name')]}", "", "| metric | value | unit |", "|---|---:|---|"]
    o2 += [f"| {k} | {row[h.index(k)]} | {units[h.index(k)]} |" for k in keys if k in h] + [""]
    print(row[h.index('Kernel Name')][:50], row[h.index('gpu__time_duration.sum')], row[h.index('dram__bytes_read.sum')],
          row[h.index('dram__bytes_write.sum')], row[h.index('smsp__inst_executed.sum')], row[h.index('smsp__issue_active.avg.pct_of_peak_sustained_active')])
open(f"profiles/{tag}_ncu_full_summary.md", "w").write("\n".join(o2))
