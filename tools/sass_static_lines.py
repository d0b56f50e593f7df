#!/usr/bin/env python
"""Static code size of one kernel by source line (no GPU needed): which source regions the SASS of a kernel comes from.
    nvdisasm -g <cubin> > all.dis ; python tools/sass_static_lines.py all.dis <mangled-name-substring> [top]"""
import re, sys
from collections import defaultdict
path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_fn, cur_line, inside = None, None, False
by_line = defaultdict(int)
by_file = defaultdict(int)
total = 0
for l in open(path, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        inside = key in m.group(1)
        cur_line = None
        continue
    if l.lstrip().startswith(".section"):
        inside = False
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        total += 1
        by_line[cur_line] += 1
        by_file[cur_line[0] if cur_line else "?"] += 1
print("instructions: %d (%.1f KB)" % (total, total * 16 / 1024))
for f, n in sorted(by_file.items(), key=lambda kv: -kv[1]):
    print("  %-30s %5d  %5.1f%%" % (f, n, 100.0 * n / max(total, 1)))
for k, n in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
    print("%-34s %5d  %5.1f%%" % ("%s:%d" % k if k else "?", n, 100.0 * n / max(total, 1)))
