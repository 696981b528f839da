"""Estimate register-operand traffic (32-bit words read from the register file / uniform file) per
warp-solve from an ncu source page: tools/operand_words.py prof.ncu-rep n_warp_solves
Rules: 64-bit ops (D*) read 2 words per register source, others 1; `.reuse` sources, RZ, immediates,
predicates cost 0; UR / c[] sources cost like registers."""
import csv, io, re, subprocess, sys, collections
rep, nsolve = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
tot_words = 0.0
by = collections.Counter(); cnt = collections.Counter()
for r in data:
    n = int(r[ia])
    if not n: continue
    text = r[isrc].strip()
    parts = text.split()
    if parts[0].startswith("@"): parts = parts[1:]
    op = parts[0]
    ops = " ".join(parts[1:]).split(",")
    wide = op.startswith(("DFMA", "DADD", "DMUL", "DSETP"))
    srcs = ops[1:] if not op.startswith(("ST", "BRA", "VOTE", "ISETP", "DSETP", "BAR", "EXIT")) else ops
    if op.startswith(("ISETP", "DSETP")): srcs = ops[2:]
    w = 0
    for o in srcs:
        o = o.strip().lstrip("-|!~").rstrip("|")
        if re.match(r"^U?R\d+", o) and "reuse" not in o: w += 2 if wide else 1
        elif o.startswith("c["): w += 2 if wide else 1
    if op.startswith("FSEL") or op.startswith("SEL"): pass
    tot_words += w * n
    by[op.split(".")[0]] += w * n; cnt[op.split(".")[0]] += n
print(f"operand words per warp-solve: {tot_words / nsolve:.0f}  -> {tot_words / nsolve / 2:.0f} cycles at 2 words/clk")
for op, w in by.most_common(14):
    print(f"  {op:8s} {w / nsolve:7.1f} words/solve  ({cnt[op] / nsolve:6.1f} instr, {w / max(1, cnt[op]):.2f} words each)")
