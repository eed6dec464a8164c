#!/usr/bin/env python
"""Join `ncu --page source --csv` (SASS view, per-instruction executed counts) with nvdisasm line info of the
in-tree libepx.so, to get executed warp-instructions per CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <kernel-regex> <mangled-substring> <units> [top] [samples]
(a 6th argument 'samples' sorts by warp-stall samples instead of executed instructions)"""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict
rep, kre, mangled, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# NCU_SKIP=<k>: take the (k+1)-th captured launch of the report instead of the first one that matches the regex
skip = os.environ.get('NCU_SKIP')
sel = ['--launch-skip', skip, '--launch-count', '1'] if skip else ['--kernel-name', 'regex:' + kre]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[h]; ia, isrc, ie, iss = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ins = []
for r in rows[h + 1:]:
    if len(r) > ie and r[ie].isdigit():
        ins.append((int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[iss]) if r[iss].isdigit() else 0))
    elif len(r) > 1 and r[0] == 'Kernel Name':
        break  # only the first matching launch
base = ins[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'epimethex_b200', 'libepx.so')], cwd=tmp, capture_output=True)
dis = ''
for f in sorted(os.listdir(tmp)):
    if f.endswith('.cubin'):
        d = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ('.text.' in d) and (mangled in d):
            dis = d
            break
cur, infn, lines = None, False, {}
for l in dis.splitlines():
    if l.startswith('.text.') or re.match(r'^\s*\.section\s+\.text\.', l):
        infn = mangled in l
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and infn:
        lines[int(m.group(1), 16)] = cur
tot = sum(i[2] for i in ins)
by = Counter(); smp = Counter(); ops = defaultdict(Counter)
for addr, src, e, s in ins:
    key = lines.get(addr - base, ('?', 0))
    by[key] += e; smp[key] += s
    ops[key][re.sub(r'^@!?U?P\d+\s+', '', src).split()[0].split('.')[0]] += e
print('total warp-instr %d = %.1f per unit; %d SASS instr; mapped %d' % (tot, tot / units, len(ins), sum(1 for a in ins if (a[0] - base) in lines)))
allsmp = sum(smp.values()) or 1
order = smp.most_common(top) if len(sys.argv) > 6 and sys.argv[6] == 'samples' else by.most_common(top)
for key, _ in order:
    e = by[key]
    print('%-22s %8.1f /unit %5.1f%%  stall-samples %5.1f%%  %s' % ('%s:%d' % key, e / units, 100.0 * e / tot, 100.0 * smp[key] / allsmp,
          ' '.join('%s:%.0f' % (o, c / units) for o, c in ops[key].most_common(5))))
