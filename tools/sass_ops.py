"""Compact opcode listing of a cuobjdump -sass dump: one mnemonic per instruction, with loop back-edges marked.
usage: python tools/sass_ops.py dump.txt [start_addr_hex end_addr_hex]"""
import re, sys
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);')
ops = []
for line in open(sys.argv[1]):
    m = pat.match(line)
    if m:
        ops.append((int(m.group(1), 16), m.group(2).strip()))
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
out = []
for a, t in ops:
    if a < lo or a > hi: continue
    toks = t.split()
    op = toks[0] if not toks[0].startswith('@') else toks[0] + ' ' + toks[1]
    if 'BRA' in op or 'BAR' in op:
        out.append('\n[%x] %s\n' % (a, t))
    else:
        out.append(op.split('.')[0] if not op.startswith('@') else op)
print(' '.join(out))
