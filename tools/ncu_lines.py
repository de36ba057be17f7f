#!/usr/bin/env python
"""Top source lines of an ncu --import-source report by warp-stall samples (first captured launch).
usage: ncu_lines.py report.ncu-rep [N]"""
import csv


def num(x, f=int):
    try:
        return f(x)
    except (ValueError, TypeError):
        return f(0)

import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
seen = set()
items = []
cur, hdr = None, None
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        if cur in seen and hdr is not None and len(seen) > 3 and cur == first:
            break  # second launch starts
        if not seen:
            first = cur
        seen.add(cur)
        hdr = None
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] == '':
        continue
    d = dict(zip(hdr, r))
    s = num(d['# Samples'])
    tot += s
    items.append((s, cur, d['Line No'], num(d['Instructions Executed']), (num(d.get('Thread Instructions Executed')) / max(1, num(d['Instructions Executed']))),
                  num(d.get('stall_long_sb')), num(d.get('stall_wait')), num(d.get('stall_math')),
                  num(d.get('stall_branch_resolving')), num(d.get('stall_short_sb')), num(d.get('stall_not_selected')), r[1].strip()[:90]))
items.sort(reverse=True)
print('total samples', tot)
print('samples  %    file:line   warp-inst  lanes | long_sb wait math branch short_sb not_sel | source')
for s, f, l, inst, thr, lsb, w, m, b, ssb, ns, src in items[:top]:
    print(f"{s:7d} {100 * s / tot:4.1f} {f}:{l} {inst:10d} {thr:4.1f} | {lsb:6d} {w:6d} {m:5d} {b:5d} {ssb:5d} {ns:5d} | {src}")
