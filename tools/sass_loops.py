"""Static schedule of every loop in a cuobjdump -sass dump: for each backward branch, the instruction mix and the sum of
the control-code stall counts of the loop body (= the minimum issue time of ONE warp for one iteration).
usage: python tools/sass_loops.py dump.txt [min_instrs]"""
import re, sys
lines = open(sys.argv[1]).read().split('\n')
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/')
pat2 = re.compile(r'^\s+/\* (0x[0-9a-f]+) \*/')
ins = []
for i, l in enumerate(lines):
    m = pat.match(l)
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            c = int(m2.group(1), 16) >> 41
            ins.append((int(m.group(1), 16), m.group(2).strip(), c & 0xf))
addr = {a: k for k, (a, _, _) in enumerate(ins)}
for k, (a, t, st) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr and k - addr[tgt] >= minlen:
            body = ins[addr[tgt]:k + 1]
            mix = {}
            for _, tt, _ in body:
                toks = tt.split(); op = toks[0] if not toks[0].startswith('@') else toks[1]
                op = op.split('.')[0]; mix[op] = mix.get(op, 0) + 1
            stall = sum(s for _, _, s in body)
            top = ', '.join('%s %d' % kv for kv in sorted(mix.items(), key=lambda kv: -kv[1])[:12])
            print('loop %05x..%05x: %d instr, sum of stall counts %d\n   %s' % (tgt, a, len(body), stall, top))
