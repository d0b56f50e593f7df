#!/usr/bin/env python
"""Registers, stack (spills), shared memory and code size of every kernel in libvanrijn_cuda.so (cuobjdump; no GPU needed).
    python tools/kernel_table.py [path/to/lib.so] [name filter regex]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 and os.path.exists(sys.argv[1]) else os.path.join(root, "vanrijn_b200/lib/libvanrijn_cuda.so")
flt = re.compile(sys.argv[-1]) if len(sys.argv) > 1 and not os.path.exists(sys.argv[-1]) else None
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
size = {}
for l in elf.splitlines():
    m = re.search(r"\.text\.(\S+)", l)
    if m and "PROGBITS" in l:
        hexs = [x for x in l.split() if re.fullmatch(r"[0-9a-f]+", x)]
        if len(hexs) > 2:
            size[m.group(1)] = int(hexs[2], 16)
rows = []
cur = None
for l in res.splitlines():
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", l)
    if m and cur:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void vrj::", "").replace("(bool)", "").replace("(int)", "")
        rows.append((name, int(m.group(1)), int(m.group(2)), int(m.group(3)), size.get(cur, 0)))
        cur = None
seen = set()
print("%-58s %5s %6s %6s %8s" % ("kernel", "regs", "stack", "smem", "code B"))
for r in sorted(rows):
    if r in seen or (flt and not flt.search(r[0])):
        continue
    seen.add(r)
    print("%-58s %5d %6d %6d %8d" % r)
