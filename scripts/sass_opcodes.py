#!/usr/bin/env python
"""Opcode histogram of the shipped kernels (cuobjdump -sass of libcgb200.so): the evidence that the
TMA bulk copy (UBLKCP + SYNCS mbarrier ops), the Philox IMAD.WIDE.U32 chains and the fp64 pipe are
what the binary contains.  Writes profiles/<tag>_sass_opcodes.md.  usage: sass_opcodes.py r02"""
import collections, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cluster_generator_b200", "lib", "libcgb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(kernels, demangle if len(demangle) == len(kernels) else list(kernels)))
# cu++filt prints template arguments as "(int)0, (bool)1": compare without the casts
names = {k: re.sub(r"\((?:int|bool)\)", "", v) for k, v in names.items()}
want = ("cgb_sample_kernel<double, 0, 0>", "cgb_gas_smem_kernel<double, 0>", "cgb_gas_kernel<double, 0>",
        "cgb_hard_kernel<double, 0, 0>", "cgb_givens_kernel<1>", "cgb_givens_kernel<0>", "cgb_eddington_g_kernel",
        "cgb_qags_replay_kernel", "cgb_screen_audit_kernel")
key_ops = ("UBLKCP", "SYNCS", "IMAD.WIDE.U32", "LOP3.LUT", "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "MUFU", "LDS", "LDG", "STG",
           "SHFL", "VOTE", "ATOM", "RED", "BAR", "WARPSYNC", "S2UR", "BRA", "CALL")
out = [f"# SASS opcode histogram of the shipped library, {tag} (`cuobjdump -sass cluster_generator_b200/lib/libcgb200.so`, sm_100a)", "",
       "Static instruction counts per kernel. `UBLKCP` = cp.async.bulk (TMA bulk copy global -> shared), `SYNCS` = mbarrier",
       "operations (init / arrive.expect_tx / try_wait); `IMAD.WIDE.U32` = the 32x32 -> 64-bit products of Philox4x32-10.", ""]
for mangled, cnt in kernels.items():
    name = names[mangled]
    if not any(w in name for w in want):
        continue
    tot = sum(cnt.values())
    out += [f"## `{name[:110]}`", "", f"{tot} instructions; top opcodes: " +
            ", ".join(f"{op} {n}" for op, n in cnt.most_common(12)), "", "| opcode family | count |", "|---|---:|"]
    for k in key_ops:
        n = sum(v for op, v in cnt.items() if op.startswith(k))
        if n:
            out.append(f"| {k}* | {n} |")
    out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.md"), "w").write("\n".join(out))
print("\n".join(out[:60]))
