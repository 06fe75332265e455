"""Developer tool: per-source-line and per-opcode dynamic instruction counts from an .ncu-rep (needs -lineinfo)."""
import csv, re, collections, subprocess, sys, io
rep, kernel, skip, iters = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
def run(extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-skip", skip, "--launch-count", "1"] + extra,
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = run([])
hdr = rows[1]; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
byop = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ie or not (r[ie] or '0').isdigit(): continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia].strip())
    op = m.group(2).split('.')[0] if m else r[ia].strip()
    n = int(r[ie] or 0); byop[op] += n; tot += n
print("kernel %s: %.1f warp-instr per warp-iteration (%d total)" % (kernel, tot / iters, tot))
fp64 = sum(byop[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "F2I", "I2F", "FRND"))
print("fp64-pipe: %.1f / iter" % (fp64 / iters))
print("  ".join("%s %.1f" % (o, n / iters) for o, n in byop.most_common(36)))
rows = run(["--print-source", "cuda,sass"])
cur = None; lines = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name') or r[0] == '': continue
    try: lines.append((cur, int(r[0]), r[1].strip(), int(r[7] or 0), int(r[6] or 0)))
    except ValueError: pass
lines.sort(key=lambda l: -l[3])
for f, ln, src, n, s in lines[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    print("%-18s %4d %7.1f/iter samp %4d  %s" % (f[:18], ln, n / iters, s, src[:105]))
