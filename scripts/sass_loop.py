#!/usr/bin/env python
"""Hot-loop SASS of an iteration kernel with per-pipe instruction counts (no GPU needed).

python scripts/sass_loop.py --match 'k_fused_waIdLi0ENS_5WaCfgILi2ELi2ELi1ELi1EEELi11E' --out profiles/sass_k_fused_wa_f64_loop.txt

Finds the longest backward-branch loop of the first function whose mangled name contains --match in
wdpm_b200/libwdpm_b200.so (cuobjdump -sass), classifies every instruction by the pipe / issue cost it has on
sm_100a as measured in round 1 (profiles/micro_chain_throughput.txt: an FP64 instruction holds the issue port
for two cycles, everything else for one) and writes the counts followed by the listing.
"""
import argparse
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ap = argparse.ArgumentParser()
ap.add_argument("--match", required=True)
ap.add_argument("--lib", default=str(ROOT / "wdpm_b200" / "libwdpm_b200.so"))
ap.add_argument("--out", default="")
ap.add_argument("--relaxes-per-iteration", type=int, default=6, help="tile relaxes one thread does per loop iteration")
a = ap.parse_args()

txt = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pick = next(f for f in funcs[1:] if a.match in f.split("\n")[0])
name = pick.split("\n")[0].strip()
ins = []
for line in pick.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, t in ins:
    m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < addr:
        loops.append((int(m.group(1), 16), addr))
def arithmetic(lo_hi):  # the compute loop is the one with the floating-point work (the copy warp's loop has none)
    return sum(1 for x, t in ins if lo_hi[0] <= x <= lo_hi[1] and re.search(r"\b(DADD|DMUL|FADD|FMUL)\b", t))


lo, hi = max(loops, key=lambda x: (arithmetic(x), x[1] - x[0]))
body = [(x, t) for x, t in ins if lo <= x <= hi]
PIPES = {"fp64": ("DADD", "DMUL", "DSETP", "DFMA"), "fp32": ("FADD", "FMUL", "FSETP", "FMNMX", "FFMA"),
         "select/int (alu)": ("FSEL", "SEL", "ISETP", "IMAD", "IADD3", "LOP3", "SHF", "LEA", "VIADD", "MOV", "VIMNMX", "PRMT", "IABS"),
         "shared memory": ("LDS", "STS"), "shuffle": ("SHFL",), "barrier/mbarrier": ("BAR", "SYNCS", "WARPSYNC", "MEMBAR", "FENCE"),
         "control": ("BRA", "BSSY", "BSYNC", "EXIT", "NOP", "CALL", "RET"), "uniform datapath": ("UMOV", "ULEA", "UIADD3", "UISETP", "S2UR", "UIMAD", "ULOP3", "USHF", "USEL", "R2UR", "LDCU")}
counts, ops = collections.Counter(), collections.Counter()
for _, t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0].split(".")[0]
    ops[op] += 1
    counts[next((p for p, names in PIPES.items() if op in names), "other")] += 1
n = len(body)
issue = n + counts["fp64"]
out = [f"kernel     {name}", f"loop       0x{lo:x} .. 0x{hi:x}: {n} instructions per iteration = one step of one thread = {a.relaxes_per_iteration} tile relaxes",
       f"issue cost {issue} cycles per warp and step (FP64 instructions count twice) = {issue / a.relaxes_per_iteration:.1f} per relax", "", "by pipe:"]
out += [f"  {p:20s} {c:5d}" for p, c in counts.most_common()]
out += ["", "by opcode:"] + [f"  {o:10s} {c:5d}" for o, c in ops.most_common()]
out += ["", "listing:"] + [f"  {x:05x}  {t}" for x, t in body]
text = "\n".join(out) + "\n"
if a.out:
    Path(a.out).write_text(text)
print("\n".join(out[:12]))
