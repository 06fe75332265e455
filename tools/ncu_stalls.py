"""Developer tool: instructions with the most warp-stall samples in an ncu source-page CSV
(ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[4].isdigit(): data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[4]) for r in data)
print("total samples", tot, "static instructions", len(data), "executed warp-instr", sum(int(r[5] or 0) for r in data))
agg = {h: sum(int(r[ci[h]] or 0) for r in data) for h in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][4]))[:top]
for i in sorted(idx):
    r = data[i]
    why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5d %-72s samp %5s exec %5s  %s" % (i, r[1].strip()[:72], r[4], r[5], why))
