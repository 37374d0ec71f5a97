#!/usr/bin/env python
"""Summarise one round's ncu captures from gpurun_out/ into profiles/ (tracked).

usage: make_profiles.py r02
       make_profiles.py summary <report.ncu-rep> <out.md> <title>   (the metric table of any capture)
inputs (written by the gpurun commands in profiles/README.md):
  gpurun_out/launches_<tag>.csv     ncu --metrics gpu__time_duration.sum launch list of bench.py
  gpurun_out/prof_<tag>.ncu-rep     ncu --set full of cgb_sample_kernel / gas kernel / f(E) kernels
  gpurun_out/bench_<tag>.json       the same bench.py command without ncu
outputs:
  profiles/<tag>_launches.{md,csv}, profiles/<tag>_ncu_full_summary.md, profiles/<tag>_bench.json,
  profiles/<tag>_ncu.json   per-launch counters of the dominant kernel + sha256 of its source;
                            bench.py reads it for roofline.traffic / roofline.issue and flags
                            "stale" when cgb_sampler.cu has changed since the capture.
"""
import collections, csv, json, os, shutil, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__warps_eligible.avg.per_cycle_active']


def summarise(report, out_md, title):
    """Selected metrics of the last capture of every kernel in an ncu report -> markdown."""
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); h, units = r[0], r[1]
    o2 = [f"# ncu --set full, {title}", "",
          "Raw report kept out of git; regenerate with the command in profiles/README.md. Values per launch (the last capture of each kernel).", ""]
    last = collections.OrderedDict((row[h.index('Kernel Name')], row) for row in r[2:])
    for row in last.values():
        o2 += [f"## {row[h.index('Kernel Name')]}", "", "| metric | value | unit |", "|---|---:|---|"]
        o2 += [f"| {k} | {row[h.index(k)]} | {units[h.index(k)]} |" for k in KEYS if k in h] + [""]
    open(out_md, "w").write("\n".join(o2))
    return h, units, last


if len(sys.argv) > 1 and sys.argv[1] == "summary":
    summarise(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else os.path.basename(sys.argv[2]))
    sys.exit(0)

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (code_sha256: the hash bench.py compares against)


def last_json(path):
    return json.loads(open(path).read().strip().split("\n")[-1])


b = last_json(f"{G}/bench_{tag}.json")
if os.path.exists(f"{G}/launches_{tag}.csv"):
    lines = [l for l in open(f"{G}/launches_{tag}.csv") if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list, {tag} (`CGB_BENCH_CONFIGS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 30 --warmup 3`)", "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES. First 600 launches = table build + f(E) builds + warm-up + 30 timed steps (+ part of the e2e leg).", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:90]}` | {a[0]} | {a[1]/1e3:.1f} | {a[1]/a[0]/1e3:.2f} | {100*a[1]/tot:.1f}% |")
    st = {k: a for k, a in agg.items() if any(x in k for x in ("cgb_sample_kernel", "cgb_gas", "cgb_hard_kernel"))}
    ts = sum(a[1] / a[0] for a in st.values())
    out += ["", "## one bench step (DM sample kernel + deferred pass + gas kernel)", "", "| kernel | avg us | share of step |", "|---|---:|---:|"]
    for k, a in st.items():
        out.append(f"| `{k[:70]}` | {a[1]/a[0]/1e3:.1f} | {100*(a[1]/a[0])/ts:.1f}% |")
    out += ["", f"bench.py (not under ncu, CUDA events): ms_per_step {b['ms_per_step']:.3f}, sample kernel "
            f"{b['roofline']['kernel_ms']:.3f} ms = {100*b['roofline']['kernel_ms']/b['ms_per_step']:.1f}% of the step, gas kernel "
            f"{b['roofline']['gas_kernel']['kernel_ms']:.3f} ms = {100*b['roofline']['gas_kernel']['kernel_ms']/b['ms_per_step']:.1f}% -- agrees with the shares above."]
    open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")
    shutil.copy(f"{G}/launches_{tag}.csv", f"profiles/{tag}_launches.csv")
json.dump(b, open(f"profiles/{tag}_bench.json", "w"))

h, units, last = summarise(f"{G}/prof_{tag}.ncu-rep", f"profiles/{tag}_ncu_full_summary.md", tag)

# counters bench.py needs, of the dominant kernel on bench.py's own workload
samp = [row for name, row in last.items() if name.startswith("void cgb_sample_kernel<double, 0, 0>") or "cgb_sample_kernel<double, 0, 0>" in name]
if samp:
    row = samp[0]
    def val(k, scale=1.0):
        u = units[h.index(k)]
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "inst": 1.0, "": 1.0}.get(u, 1.0)
        return float(row[h.index(k)].replace(",", "")) * mult * scale
    prof = {
        "kernel": row[h.index('Kernel Name')],
        "workload": b["config"]["workload"], "n_particles": b["config"]["dm_per_gpu"],
        "warp_instructions_per_launch": val('smsp__inst_executed.sum'),
        "dram_bytes_per_launch": val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
        "gpu_time_us_under_ncu": val('gpu__time_duration.sum'),
        "sampler_sha256": bench.code_sha256("cluster_generator_b200/csrc/cgb_sampler.cu"),
        "source": f"gpurun_out/prof_{tag}.ncu-rep (ncu --set full --clock-control none), summarised by scripts/make_profiles.py",
    }
    json.dump(prof, open(f"profiles/{tag}_ncu.json", "w"), indent=1)
    print(prof)
