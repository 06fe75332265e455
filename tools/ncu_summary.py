"""Developer tool: selected raw metrics of every kernel in an .ncu-rep (ncu --set full) as JSON, the form committed
under profiles/.  usage: python tools/ncu_summary.py X.ncu-rep > profiles/NAME_ncu_full_summary.json"""
import csv, io, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
REASONS = ["wait", "long_scoreboard", "not_selected", "selected", "short_scoreboard", "barrier", "math_pipe_throttle",
           "branch_resolving", "dispatch_stall", "no_instruction", "mio_throttle", "lg_throttle", "membar", "drain", "imc_miss"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    if len(r) != len(hdr): continue
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    k = name if name not in res else "%s #%s" % (name, d["ID"])
    e = {m: ("%s %s" % (d[m], u[m])).strip() for m in WANT if m in d}
    st = {}
    for s in REASONS:
        m = STALL % s
        if m in d and d[m] not in ("", "n/a"):
            try: st[s] = round(float(d[m].replace(",", "")), 3)
            except ValueError: pass
    e["stall_cycles_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:9])
    res[k] = e
print(json.dumps(res, indent=1))
