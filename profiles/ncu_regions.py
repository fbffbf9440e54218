#!/usr/bin/env python
"""Stall reasons per code region of the rows kernel: ncu -i X.ncu-rep --page source --csv | python ncu_regions.py"""
import csv, collections, sys
rows = list(csv.reader(sys.stdin))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hdr]; body = rows[hdr + 1:]
iS, iN, iX = 2, 1, 5
reasons = [(i, h) for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
def op(r):
    t = r[iN].split()
    return t[1] if t[0].startswith('@') else t[0]
N = len(body)
tot = sum(int(r[iS] or 0) for r in body)
regions = collections.defaultdict(collections.Counter)
for i, r in enumerate(body):
    ops = [op(body[j]) for j in range(max(0, i - 25), min(N, i + 25))]
    nf = sum(o.startswith('FFMA') or o.startswith('FMUL') for o in ops)
    ns = sum(o.startswith('STG') or o.startswith('ST.E') for o in ops)
    nsy = sum('SYNCS' in o for o in ops)
    reg = 'rows' if nf >= 15 else 'out' if ns >= 6 else 'sync/wait' if nsy >= 1 else 'other'
    for k, h in reasons:
        regions[reg][h] += int(r[k] or 0)
    regions[reg]['_instr'] += int(r[iX] or 0)
for reg, c in regions.items():
    t = sum(v for k, v in c.items() if k != '_instr')
    print(f"== {reg}: samples {100 * t / tot:.1f}%  instr {c['_instr'] / 1e6:.1f}M")
    for k, v in c.most_common(8):
        if k != '_instr' and v > 0.02 * t:
            print(f"      {k:28s} {100 * v / t:5.1f}%")
