"""Developer tool: dynamic warp-instructions and stall samples per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file` (needs -lineinfo).
usage: ncu_bylines.py file kernel-substring warp_iterations [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want, iters = sys.argv[2], float(sys.argv[3]); top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
cur_file = cur_fn = None; hdr = None
per = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) == 2 and r[0] == 'Function Name': cur_fn = r[1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or cur_fn is None or want not in cur_fn or len(r) != len(hdr): continue
    try: ln = int(r[0])
    except ValueError: continue
    ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    k = (cur_file, ln)
    e = per.setdefault(k, [r[1].strip(), 0, 0])
    e[1] += int(r[ie] or 0); e[2] += int(r[isamp] or 0)
tot = sum(v[1] for v in per.values()); ts = sum(v[2] for v in per.values())
print("%s: %.1f warp-instr per warp-iteration, %d samples" % (want, tot / iters, ts))
byfile = collections.Counter()
for (f, ln), v in per.items(): byfile[f] += v[1]
print({f: round(n / iters, 1) for f, n in byfile.items()})
for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-16s %4d %7.1f/iter  samp %5.1f%%  %s" % (f[:16], ln, v[1] / iters, 100.0 * v[2] / max(ts, 1), v[0][:110]))
