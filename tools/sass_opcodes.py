#!/usr/bin/env python
"""SASS opcode histogram of the step's kernels (cuobjdump -sass; no GPU needed) -> profiles/rNN_sass_opcodes.txt.
Shows what the kernels are made of: binary64 arithmetic (DADD / DMUL / DFMA), 256-bit read-only loads (LDG.E.256 ...), and
that there are no tensor-core (HMMA / UTCMMA ...) or TMA (UBLKCP ...) opcodes on this path (DESIGN 3.2 says why)."""
import os, re, subprocess, sys
from collections import Counter
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "vanrijn_b200/lib/libvanrijn_cuda.so")
want = ["k_raygen<double, false>", "k_trace_rec<false>", "k_shade<float, double, false, false, true, 1>",
        "k_shade<float, double, false, false, false, 1>", "k_shade<float, double, false, false, false, 15>", "k_resolve<double>",
        "k_tail<float, double, false, false, 1>", "k_shade<float, double, false, true, false, 15>"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts, wide = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void vrj::", "").replace("(bool)", "").replace("(int)", "")
        cur = None if name in counts else name  # the same instantiation can sit in several object files: count it once
        if cur:
            counts[cur], wide[cur] = Counter(), Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        if m.group(1) in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOMG", "RED", "ATOM", "LDGSTS"):
            wide[cur][m.group(1) + m.group(2)] += 1
def norm(n):
    return n.replace("(bool)0", "false")
for w in want:
    key = None
    for k in counts:
        kk = k.replace(" ", "")
        if kk == w.replace(" ", "").replace("false", "0").replace("true", "1") or kk == w.replace(" ", ""):
            key = k
    if key is None:
        cands = [k for k in counts if k.replace("false", "0").replace("true", "1").replace(" ", "") == w.replace("false", "0").replace("true", "1").replace(" ", "")]
        key = cands[0] if cands else None
    if key is None:
        print("## %s: not found" % w)
        continue
    c = counts[key]
    total = sum(c.values())
    print("## %s: %d instructions (%.1f KB)" % (w, total, total * 16 / 1024))
    print("   " + ", ".join("%s %d" % kv for kv in c.most_common(24)))
    print("   memory: " + ", ".join("%s %d" % kv for kv in wide[key].most_common(12)))
    tensor = [o for o in c if re.match(r"(HMMA|IMMA|DMMA|UTC|QGMMA|WGMMA|UBLKCP|UTMA|SYNCS)", o)]
    print("   tensor / TMA opcodes: %s" % (", ".join(tensor) if tensor else "none"))
