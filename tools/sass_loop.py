"""Developer tool: static opcode histogram of the largest loop of a kernel (cuobjdump -sass text on stdin).
usage: cuobjdump -sass -fun <mangled> lib.so | python tools/sass_loop.py"""
import re, sys, collections
ins = []
for l in sys.stdin:
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA\S*\s+.*?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            span = i - addr[tgt]
            if best is None or span > best[0]: best = (span, addr[tgt], i)
print("instructions in function:", len(ins))
if best:
    span, lo, hi = best
    body = ins[lo:hi + 1]
    ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in body)
    print("largest loop: %d instructions (0x%x..0x%x)" % (len(body), body[0][0], body[-1][0]))
    print("  ".join("%s %d" % kv for kv in ops.most_common(40)))
    f64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "F2I", "I2F", "FRND", "MUFU"))
    print("fp64-pipe-ish:", f64, " LDL/STL:", ops["LDL"] + ops["STL"], " BRA:", ops["BRA"], " CALL:", ops["CALL"])
