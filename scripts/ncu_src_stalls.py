#!/usr/bin/env python
"""Stall samples per SASS instruction from `ncu --page source --csv` (usage: ncu_src_stalls.py src.csv lo hi | utc)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
if sys.argv[2] == 'utc':
    for i, r in enumerate(data):
        if 'UTCHMMA' in r[ix['Source']] or 'UTCBAR' in r[ix['Source']] or 'UTMALDG' in r[ix['Source']]:
            print(i, r[ix['Source']][:60], int(g(r, 'Instructions Executed')), int(g(r, '# Samples')))
    sys.exit()
lo, hi = int(sys.argv[2]), int(sys.argv[3])
tot = 0
for i, r in enumerate(data[lo:hi]):
    s = g(r, '# Samples'); tot += s
    top = sorted([(g(r, c), c) for c in stall_cols], reverse=True)[:3]
    print(lo + i, "%-62s" % r[ix['Source']][:62], int(s), int(g(r, 'Instructions Executed')), [(c[6:], int(v)) for v, c in top if v > 0])
print("total samples", tot)
