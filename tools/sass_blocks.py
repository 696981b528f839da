"""Static instruction count of one rv_lnl_kernel build, per basic block.

    python tools/sass_blocks.py [lib.so] [U] [THREADS] [lo_hex hi_hex]

Disassembles the library with cuobjdump, keeps the instantiation rv_lnl_kernel<0, U, THREADS>,
cuts the address range [lo, hi] (default: everything) into basic blocks at branch targets and
branches, and prints FP64 (DFMA/DMUL/DADD/DSETP) and other instructions per block with the
opcode mix of the others.  Weighted by the passes per solve (bench key mean_newton_iters, the
histogram of the per-warp step maximum) this predicts a change before it is measured: an FP64
instruction costs two issue cycles of a sub-partition, any other one.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "evidence_b200", "librvlnl.so")
U = sys.argv[2] if len(sys.argv) > 2 else "4"
T = sys.argv[3] if len(sys.argv) > 3 else "512"
lo = int(sys.argv[4], 16) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5], 16) if len(sys.argv) > 5 else 1 << 30

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
want = f"rv_lnl_kernelILi0ELi{U}ELi{T}E"
ins = []
keep = False
for line in sass.splitlines():
    if "Function :" in line:
        keep = want in line
        continue
    if not keep:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\*", line)
    if m and lo <= int(m.group(1), 16) <= hi:
        ins.append((int(m.group(1), 16), m.group(2)))
if not ins:
    sys.exit(f"no instructions of {want} in {lib}")

targets = set()
for _, text in ins:
    m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
    if m:
        targets.add(int(m.group(1), 16))

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
blocks, cur = [], None
for addr, text in ins:
    if addr in targets and cur:
        blocks.append(cur)
        cur = None
    if cur is None:
        cur = dict(start=addr, end=addr, F=0, O=0, ops={}, last="")
    words = text.split()
    op = (words[1] if words[0].startswith("@") else words[0]).split(".")[0]
    if op in FP64:
        cur["F"] += 1
    else:
        cur["O"] += 1
        cur["ops"][op] = cur["ops"].get(op, 0) + 1
    cur["end"], cur["last"] = addr, text
    if "BRA" in text or "EXIT" in text or "RET" in text:
        blocks.append(cur)
        cur = None
if cur:
    blocks.append(cur)
for b in blocks:
    mix = " ".join(f"{k}:{v}" for k, v in sorted(b["ops"].items(), key=lambda kv: -kv[1]))
    print(f"{b['start']:05x}-{b['end']:05x} F={b['F']:3d} O={b['O']:3d}  {mix}   | {b['last'][:44]}")
