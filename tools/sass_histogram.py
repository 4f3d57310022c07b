"""Opcode histogram per kernel of libsgp.so (cuobjdump -sass): what the shipped binary executes.  usage: sass_histogram.py [libsgp.so] > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaussianprocessnode_b200", "libsgp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "LD", "ST", "RED", "ATOM", "ATOMG", "SHFL", "BAR", "MEMBAR", "CCTL", "LDSM", "HMMA", "UTCHMMA"]
name = None; cnt = collections.Counter(); res = []
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        if name: res.append((name, cnt))
        name = m.group(1); cnt = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        cnt[m.group(1)] += 1; cnt["_total"] += 1
if name: res.append((name, cnt))
demangle = subprocess.run(["c++filt"], input="\n".join(n for n, _ in res), capture_output=True, text=True).stdout.splitlines()
print("SASS opcode histogram of %s (sm_100a)\n" % os.path.basename(lib))
for (n, c), d in sorted(zip(res, demangle), key=lambda t: -t[0][1]["_total"]):
    short = re.sub(r"\(anonymous namespace\)::", "", d)[:150]
    print("%-150s  %6d instr | %s" % (short, c["_total"], "  ".join("%s %d" % (k, c[k]) for k in KEY if c[k])))
