#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm -g line info to get per-source-line instruction counts.

usage: sass_lines.py <ncu_source_page.csv> <nvdisasm -g -c output> <kernel substring> [top N]
"""
import csv
import re
import sys

ncu_csv, sass, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ncu = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]] or 0)) for r in rows[2:] if len(r) > 5]
# nvdisasm: find function section, collect (file,line) per instruction in order
lines, cur, infn = [], ("?", 0), False
for l in open(sass):
    if l.startswith(".text.") or "\t.section\t.text." in l or l.strip().startswith(".section"):
        infn = kern in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
n = min(len(lines), len(ncu))
print(f"# {len(ncu)} ncu rows, {len(lines)} disassembled instructions")
agg = {}
for i in range(n):
    k = lines[i]
    a = agg.setdefault(k, [0, 0])
    a[0] += ncu[i][1]
    a[1] += ncu[i][2]
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"# total warp-instructions {tot}, samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[1]/tots:5.1f}% samples  {k[0]}:{k[1]}")
