"""Summarise an ncu report of rv_lnl_kernel: key metrics, instruction mix, stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_warp_solves]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
nsolve = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ia, isrc = h.index("Instructions Executed"), h.index("Source")
byop = collections.Counter()
for r in data:
    parts = r[isrc].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    byop[op.split(".")[0]] += int(r[ia])
tot = sum(byop.values())
fp64 = sum(byop[o] for o in ("DFMA", "DADD", "DMUL", "DSETP"))
print(f"\nwarp-instructions {tot:.4g}; FP64 {fp64:.4g} ({100 * fp64 / tot:.1f}%)")
if nsolve:
    print(f"per warp-solve: total {tot / nsolve:.1f}  fp64 {fp64 / nsolve:.1f}  other {(tot - fp64) / nsolve:.1f}  "
          f"cycles/SMSP {float(m['smsp__cycles_active.avg']) * 592 / nsolve:.1f}")
for op, c in byop.most_common(24):
    print(f"  {op:10s} {100 * c / tot:5.1f}%" + (f"  {c / nsolve:6.1f}/solve" if nsolve else ""))
st = collections.Counter()
for i, name in enumerate(h):
    if name.startswith("stall_") and "Not Issued" not in name:
        st[name] = sum(int(r[i]) for r in data)
ts = sum(st.values())
print()
for k, v in st.most_common(8):
    print(f"  {k:26s} {100 * v / ts:5.1f}%")
