"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_by_line.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit() and r[2] == '-':   # a source line row (aggregated)
        d = dict(zip(hdr, r))
        def gi(k):
            try: return int(d[k])
            except: return 0
        lines.append((gi('# Samples'), gi('Instructions Executed'), cur, int(r[0]), r[1].strip(),
                      gi('stall_long_sb'), gi('stall_no_inst'), gi('stall_wait'), gi('stall_short_sb'), gi('stall_math'),gi('stall_lg'), gi('stall_branch_resolving')))
tot = sum(l[0] for l in lines) or 1; toti = sum(l[1] for l in lines) or 1
print('total samples', tot, 'warp instructions', toti)
byfile = {}
for l in lines:
    byfile.setdefault(l[2], [0,0]); byfile[l[2]][0]+=l[0]; byfile[l[2]][1]+=l[1]
for f,(a,b) in byfile.items(): print(f'{f:20s} samples {100*a/tot:5.1f}%  inst {100*b/toti:5.1f}%')
print(' %samp  %inst  long_sb no_inst wait   file:line  source')
for l in sorted(lines, reverse=True)[:top]:
    print(f'{100*l[0]/tot:5.1f}% {100*l[1]/toti:5.1f}%  {100*l[5]/max(l[0],1):4.0f}% {100*l[6]/max(l[0],1):4.0f}% {100*l[7]/max(l[0],1):4.0f}%  {l[2]}:{l[3]}  {l[4][:100]}')
