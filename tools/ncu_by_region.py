"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by code region (function-level line ranges).
Usage: python tools/ncu_by_region.py src.csv"""
import csv, sys, re, os
here = os.path.dirname(os.path.abspath(__file__))
def regions(path):
    out = []
    for n, l in enumerate(open(path), 1):
        m = re.match(r'^(?:__device__|DM_API|DM_HD|__global__|static|template).*?\b([A-Za-z_0-9]+)\s*\(', l)
        if m and not l.startswith('  '): out.append((n, m.group(1)))
    return out
files = {f: regions(os.path.join(here, '..', 'samsim_b200', 'csrc', f)) for f in ['physics.cuh', 'step.cuh', 'detmath.h', 'samsim_b200.cu']}
def region_of(f, line):
    name = '?'
    for n, nm in files.get(f, []):
        if n <= line: name = nm
        else: break
    return name
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == 'Function Name': continue
    if r[0].isdigit() and r[2] == '-':
        d = dict(zip(hdr, r))
        gi = lambda k: int(d[k]) if d.get(k, '').lstrip('-').isdigit() else 0
        key = (cur, region_of(cur, int(r[0])))
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += gi('# Samples'); a[1] += gi('Instructions Executed'); a[2] += gi('stall_long_sb'); a[3] += gi('stall_wait')
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1
print(' %samp  %inst  long_sb  wait  region')
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:30]:
    print(f'{100*a[0]/ts:5.1f}% {100*a[1]/ti:5.1f}%  {100*a[2]/max(a[0],1):4.0f}%  {100*a[3]/max(a[0],1):4.0f}%  {k[0]}:{k[1]}')
