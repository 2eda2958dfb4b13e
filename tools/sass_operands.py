#!/usr/bin/env python
"""Classifies the FP instructions of a kernel's hot loop by where their source operands come from.

    python tools/sass_operands.py [libmvrl.so] [mangled-kernel-name-regex]

On sm_100 an FFMA2 / FMUL2 / FADD2 holds the FMA pipe for 2 cycles when at most two of its sources are vector
registers (immediates, uniform registers and constants are free) and for 3 cycles when all three are
(tools/ffma_regs.cu: 0.333 instead of 0.5 warp-inst/clk/SMSP).  The script finds the largest backward branch of the
kernel (the RK4 sub-step loop), drops the rare-path blocks inside it (forward branches over more than 64
instructions) and prints the instruction mix, the number of three-register packed instructions and the resulting
lower bound on FMA-pipe cycles per trip - the figure DESIGN.md section 4 quotes.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "marinevehiclereinforcementlearning_b200", "libmvrl.so")
pat = sys.argv[2] if len(sys.argv) > 2 else r"rov6_step_kernelINS_2F2ELi0ELb1ELb0"

names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout  # any listing naming the functions
fn = sorted(set(re.findall(r"_ZN4mvrl[\w]*" + pat + r"[\w]*", names)), key=len)
if not fn:
    sys.exit("no kernel matches " + pat)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], lib], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
lo = hi = None
for a, t in ins:
    m = re.search(r"BRA(?:\.U)? !?U?P\d+, 0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and (hi is None or a - int(m.group(1), 16) > hi - lo):
        lo, hi = int(m.group(1), 16), a
rare = []
for a, t in ins:
    m = re.search(r"@!?P\d BRA 0x([0-9a-f]+)", t)
    if m and lo < a < hi and int(m.group(1), 16) - a > 0x400:
        rare.append((a + 0x10, int(m.group(1), 16)))
loop = [(a, t) for a, t in ins if lo <= a <= hi and not any(x <= a < y for x, y in rare)]
mix, packed, cycles, three = collections.Counter(), collections.Counter(), 0, 0
for a, t in loop:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0].split(".")[0]
    mix[op] += 1
    if op in ("FFMA2", "FMUL2", "FADD2"):
        srcs = [o.strip() for o in t.split(None, 1)[1].split(",")[1:]]
        nreg = sum(1 for o in srcs if re.match(r"^[-|]*R\d+", o))
        packed[(op, nreg)] += 1
        cycles += 3 if nreg == 3 else 2
        three += nreg == 3
    elif op in ("FFMA", "FMUL", "FADD", "IMAD"):
        cycles += 1
print(fn[0])
print("loop 0x%x..0x%x: %d instructions per trip (%d rare-path blocks skipped)" % (lo, hi, len(loop), len(rare)))
print("mix:", ", ".join("%s %d" % kv for kv in mix.most_common(14)))
print("packed FP by number of vector-register sources:", ", ".join("%s/%d: %d" % (k[0], k[1], v) for k, v in sorted(packed.items())))
print("three-register packed instructions: %d;  FMA-pipe cycles per trip >= %d" % (three, cycles))
