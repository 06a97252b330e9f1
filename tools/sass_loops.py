"""Static per-loop instruction mix of one device function in the step kernel's SASS (nvdisasm -g -c of the cubin).
Usage: python tools/sass_loops.py <disassembly.sass> <function-substring>    (not part of the product)"""
import collections, re, sys
lines = open(sys.argv[1]).read().split('\n')
name = sys.argv[2]
i0 = next(i for i, l in enumerate(lines) if l.startswith('$') and name in l and l.rstrip().endswith(':'))
i1 = next(i for i in range(i0 + 1, len(lines)) if lines[i].lstrip().startswith('.type') or lines[i].startswith('//-----'))
inst, cur = [], None
for l in lines[i0:i1]:
    m = re.search(r'(step|physics)\.cuh", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2)))
    m3 = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);', l)
    if m3: inst.append((int(m3.group(1), 16), m3.group(2).strip(), cur))
    m4 = re.match(r'(\.L_x_\d+):', l)
    if m4: inst.append((None, m4.group(1), cur))
pos = {t: i for i, (a, t, c) in enumerate(inst) if a is None}
print(name, 'instructions', sum(1 for a, _, _ in inst if a is not None))
for idx, (a, t, cl) in enumerate(inst):
    m = re.search(r'BRA\s+(?:!?U?P\d+,\s*)?`\((\.L_x_\d+)\)', t)
    if a is not None and m and m.group(1) in pos and pos[m.group(1)] < idx:
        seg = [(tt, c) for (ad, tt, c) in inst[pos[m.group(1)]:idx + 1] if ad is not None]
        ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', tt).split()[0].split('.')[0] for tt, _ in seg)
        src = sorted({c for _, c in seg if c})
        lo = [c for c in src if c[0] == 'step']
        print(f"loop ninst {len(seg):5d} step.cuh lines {lo[0][1] if lo else '?'}-{lo[-1][1] if lo else '?'}: LDL {ops['LDL']} STL {ops['STL']} LDG {ops['LDG']+ops['LD']} STG {ops['STG']+ops['ST']} "
              f"DFMA {ops['DFMA']} DMUL {ops['DMUL']} DADD {ops['DADD']} MUFU {ops['MUFU']} CALL {ops['CALL']} BRA {ops['BRA']}")
