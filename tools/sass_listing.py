"""Developer tool: static SASS evidence of the hot kernels of libmsgwam_b200.so -- per kernel the registers, the size of
the function and of its largest loop (the ray sweep), the opcode histogram of that loop and the mnemonics that show what
the source claims (native shared-memory integer atomics, integer warp reductions, TMA bulk copy, fp64 global reductions,
the fast division / square-root seeds).   usage: python tools/sass_listing.py > profiles/r02_sass_hot_loops.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "python-msgwam_b200", "msgwam_b200", "libmsgwam_b200.so")
text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for l in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
WANT = [("column_pass_nzILi0ELb0ELb0E", "column_pass_nz<0> (N(z) pass A, one GPU)"), ("column_pass_nzILi1ELb0ELb0E", "column_pass_nz<1> (N(z) pass B)"),
        ("column_pass_nzILi1ELb0ELb1E", "column_pass_nz<1, CLAMP> (pass B with the fused post-step clamp)"),
        ("column_passILi0ELi1ELi768ELb1ELb0ELb0E", "column_pass<0> (constant N pass A, fused, one GPU)"),
        ("column_passILi1ELi1ELi768ELb1ELb0ELb0E", "column_pass<1> (constant N pass B)"),
        ("column_passILi1ELi1ELi768ELb1ELb1ELb0E", "column_pass<1, P2P> (pass B with the peer-memory exchange)"),
        ("column_frozenILb0E", "column_frozen (extension M2, single sweep)"), ("stage_rays_kernel", "stage_rays_kernel (general modes)"),
        ("scatter_kernel", "scatter_kernel (ray deletion)")]
EVID = ["ATOMS.ADD", "ATOMS.CAST.SPIN", "ATOMS.CAS", "REDUX.SUM", "CREDUX", "UBLKCP", "SYNCS", "REDG.E.ADD.F64", "ATOMG.E.ADD.F64", "MUFU.RCP64H",
        "MUFU.RSQ64H", "F2I.S64.F64", "F2F.F32.F64", "LDL", "STL", "ST.E.128.STRONG.SYS", "LD.E.128.STRONG.SYS", "CCTL"]
print("# static SASS of the hot kernels (sm_100a cubin inside libmsgwam_b200.so); made by tools/sass_listing.py")
for key, title in WANT:
    hit = [k for k in funcs if key in k]
    if not hit:
        print("\n## %s: not found" % title); continue
    ins = funcs[hit[0]]
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr and (best is None or i - addr[tgt] > best[0]):
                best = (i - addr[tgt], addr[tgt], i)
    body = ins[best[1]:best[2] + 1] if best else ins
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    hist = collections.Counter(op(t).split(".")[0] for _, t in body)
    print("\n## %s\n%s" % (title, hit[0]))
    print("instructions: %d in the function, %d in its largest loop (static; lanes execute a fraction of the branches)" % (len(ins), len(body)))
    print("loop opcodes: " + "  ".join("%s %d" % kv for kv in hist.most_common(28)))
    full = collections.Counter()
    for _, t in ins:
        o = op(t)
        for e in EVID:
            if o.startswith(e) or e in t.split(" ")[0:2] and o.startswith(e.split(".")[0]) and e in t:
                full[e] += 1
    print("evidence (whole function): " + ("  ".join("%s x%d" % (e, full[e]) for e in EVID if full[e]) or "-"))
