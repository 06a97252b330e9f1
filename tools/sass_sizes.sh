#!/bin/bash
# SASS size (bytes) of every device function in the built library, largest first
LIB=${1:-samsim_b200/_lib/libsamsim_b200.so}
cuobjdump -elf "$LIB" 2>/dev/null | awk '/^\.section \.symtab/{c++} c==2' | awk '/^\.section/{n++} n<=1' | grep -v "^\.section\|index" | python -c "
import sys
out=[]
for l in sys.stdin:
    p=l.split()
    if len(p)<7: continue
    try: sz=int(p[2],16) if p[2].startswith('0x') else int(p[2])
    except: continue
    out.append((sz,p[-1].replace('\$_Z18samsim_step_kernel7KParams\$','  step:')[:90]))
for s,n in sorted(out,reverse=True)[:int('${2:-30}')]: print(s,n)
"
