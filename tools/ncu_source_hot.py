#!/usr/bin/env python
"""Per-source-line hot spots of one profiled launch (needs a `--set full --import-source on` report and -lineinfo).

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > k.csv
    python tools/ncu_source_hot.py k.csv [top]

Prints, per (file, line): warp instructions executed, their share, stall samples, the dominant stall reasons.
"""
import csv
import os
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    cur_file, header = None, None
    inst = defaultdict(float)
    samples = defaultdict(float)
    stalls = defaultdict(lambda: defaultdict(float))
    text = {}
    opc = defaultdict(lambda: defaultdict(float))
    cur_line = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
            header = None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            ix = {n: i for i, n in enumerate(header)}
            # the header has two "Source" columns (CUDA text, SASS text)
            i_inst = ix["Instructions Executed"]
            i_samp = ix["# Samples"]
            stall_cols = [(n, i) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n]
            continue
        if header is None:
            continue
        if r[0] != "":
            cur_line = (cur_file, int(r[0]))
            text[cur_line] = r[1].strip()
            continue
        # SASS row belonging to cur_line
        try:
            n = float(r[i_inst])
        except (ValueError, IndexError):
            continue
        inst[cur_line] += n
        samples[cur_line] += float(r[i_samp] or 0)
        op = r[3].split()[0] if r[3].split() else "?"
        if op.startswith("@"):
            op = r[3].split()[1]
        opc[cur_line][op.split(".")[0]] += n
        for name, i in stall_cols:
            try:
                stalls[cur_line][name] += float(r[i] or 0)
            except ValueError:
                pass
    total = sum(inst.values())
    tot_s = sum(samples.values())
    print("total warp instructions %.0f, stall samples %.0f" % (total, tot_s))
    by_file = defaultdict(float)
    for k, v in inst.items():
        by_file[k[0]] += v
    for f, v in sorted(by_file.items(), key=lambda kv: -kv[1]):
        print("  %-28s %5.1f %%" % (f, 100 * v / total))
    print("%-26s %7s %7s  %-34s %s" % ("file:line", "inst%", "samp%", "top stalls", "source"))
    for k, v in sorted(inst.items(), key=lambda kv: -max(kv[1] / total, samples[kv[0]] / max(tot_s, 1)))[:top]:
        st = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:2]
        sts = " ".join("%s=%.0f%%" % (n.replace("stall_", ""), 100 * x / max(samples[k], 1)) for n, x in st)
        print("%-26s %6.2f%% %6.2f%%  %-34s %s" % ("%s:%d" % k, 100 * v / total, 100 * samples[k] / max(tot_s, 1), sts, text[k][:90]))
    # opcode mix
    mix = defaultdict(float)
    for k in opc:
        for o, n in opc[k].items():
            mix[o] += n
    print("opcode mix:", ", ".join("%s %.1f%%" % (o, 100 * n / total) for o, n in sorted(mix.items(), key=lambda kv: -kv[1])[:18]))


if __name__ == "__main__":
    main()
