#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples: ncu -i X.ncu-rep --page source --csv | python ncu_hot.py [N]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hdr]
iS, iN, iX = H.index('Warp Stall Sampling (All Samples)'), H.index('Source'), H.index('Instructions Executed')
body = rows[hdr + 1:]
tot = sum(int(r[iS] or 0) for r in body)
totx = sum(int(r[iX] or 0) for r in body)
print('total samples', tot, 'instructions executed', totx)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
order = sorted(range(len(body)), key=lambda i: -int(body[i][iS] or 0))[:n]
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {int(r[iS] or 0):7d} {100*int(r[iS] or 0)/tot:5.1f}%  x{int(r[iX] or 0):9d}  {r[iN].strip()[:100]}")
