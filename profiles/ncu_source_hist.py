#!/usr/bin/env python
"""Per-opcode / per-region executed-instruction histogram from `ncu --page source --csv` output.
usage: ncu_source_hist.py file.csv units_per_launch [bucket]"""
import csv, re, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); B = int(sys.argv[3]) if len(sys.argv) > 3 else 400
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[h]; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
data = []
for r in rows[h + 1:]:
    if len(r) > ie and r[ie].isdigit():
        data.append((r[ia].strip(), int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0))
tot = sum(d[1] for d in data)
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "")
print("total warp instr %d  per unit %.1f  sass lines %d" % (tot, tot / units, len(data)))
byop = Counter()
for s, e, _ in data:
    op = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
    byop[op] += e
print("per unit by opcode:", [(k, round(v / units, 1)) for k, v in byop.most_common(24)])
for i in range(0, len(data), B):
    e = sum(d[1] for d in data[i:i + B]); sm = sum(d[2] for d in data[i:i + B])
    print("%6d  %8.1f instr/unit  %7d samples | %s" % (i, e / units, sm, data[i][0][:60]))
