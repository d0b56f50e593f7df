#!/usr/bin/env python
"""Per-kernel registers / stack / spills from ptxas -v (no GPU needed)."""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = sys.argv[1:]
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
                      "--extended-lambda", "-Xptxas", "-v", "-c", "-o", "/tmp/vrj_regs.o"] + extra +
                     [os.path.join(root, "vanrijn_b200/csrc/vanrijn_cuda.cu")], capture_output=True, text=True).stderr
cur = None
rows = {}
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void vrj::", "")
        rows[cur] = {}
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and "stack" not in rows[cur]:
        rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        rows[cur]["regs"] = int(m.group(1))
for k, v in rows.items():
    print("%-48s regs %3d  stack %4d  spill st/ld %4d/%4d" % (k, v.get("regs", -1), v.get("stack", 0), v.get("spill_st", 0), v.get("spill_ld", 0)))
