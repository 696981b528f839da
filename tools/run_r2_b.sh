# round 2, GPU call B (1 GPU): high-e diagnostic + the lean Newton loop A/B
set -x
python tests/diag/diag_highecc.py 3 2>&1 | tail -8
RVL_LIB=evidence_b200/variants/librvlnl_r1.so python tests/diag/diag_highecc.py 3 2>&1 | tail -6 | sed "s/^/[r1] /"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for ilp in 2 3 4; do
  python tools/prof_sweep.py 3 131072 $ilp 2>&1 | tail -1 | sed "s/^/[lean] /"
  python tools/prof_sweep.py 2 4096 $ilp 2>&1 | tail -1 | sed "s/^/[lean] /"
done
python tools/prof_sweep.py 3 131072 2 32 2>&1 | tail -1 | sed "s/^/[lean w32] /"
python tools/prof_sweep.py 3 131072 2 24 2>&1 | tail -1 | sed "s/^/[lean w24] /"
python tools/prof_sweep.py 3 131072 2 20 2>&1 | tail -1 | sed "s/^/[lean w20] /"
RVL_LIB=evidence_b200/variants/librvlnl_r1.so python tools/prof_sweep.py 3 131072 2 2>&1 | tail -1 | sed "s/^/[r1] /"
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2b_c3 python tools/prof_sweep.py 3 131072 > gpurun_out/r2b_ncu3.log 2>&1
ls -la gpurun_out/prof_r2b_c3.ncu-rep
