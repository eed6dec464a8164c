"""Print the per-launch table of an `ncu --csv --metrics ...` log: python profiles/launch_table.py <csv> [max_rows]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid], r[ik]), {})[r[im]] = r[iv]
for (i, k), m in list(d.items())[: int(sys.argv[2]) if len(sys.argv) > 2 else 1000]:
    t = float(m.get("gpu__time_duration.sum", "nan").replace(",", ""))
    rest = {a.split(".")[0].split("__")[-1]: b for a, b in m.items() if a != "gpu__time_duration.sum"}
    print(f"{i:>4} {k.split('(')[0][:34]:34s} {t / 1e6:9.4f} ms  {rest}")
