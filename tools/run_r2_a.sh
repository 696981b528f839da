# round 2, GPU call A (1 GPU): tests, smoke, the re-anchored bench, kernel A/B sweeps, one ncu capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench1.log 2> gpurun_out/r2a_bench1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench1.err
for v in r1 no_int_tol no_final no_fma_f; do
  RVL_LIB=evidence_b200/variants/librvlnl_$v.so python tools/prof_sweep.py 3 131072 2>&1 | tail -1 | sed "s/^/[$v] /"
  RVL_LIB=evidence_b200/variants/librvlnl_$v.so python tools/prof_sweep.py 2 4096 2>&1 | tail -1 | sed "s/^/[$v] /"
done
for ilp in 2 3 4; do
  python tools/prof_sweep.py 3 131072 $ilp 2>&1 | tail -1 | sed "s/^/[new] /"
  python tools/prof_sweep.py 2 4096 $ilp 2>&1 | tail -1 | sed "s/^/[new] /"
done
python tools/prof_sweep.py 3 131072 2 24 2>&1 | tail -1 | sed "s/^/[new w24] /"
python tools/prof_sweep.py 3 131072 2 32 2>&1 | tail -1 | sed "s/^/[new w32] /"
python tools/prof_sweep.py 3 131072 3 16 2>&1 | tail -1 | sed "s/^/[new u3 w16] /"
python tools/prof_sweep.py 3 131072 4 12 2>&1 | tail -1 | sed "s/^/[new u4 w12] /"
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2a_c3 python tools/prof_sweep.py 3 131072 > gpurun_out/r2a_ncu3.log 2>&1
ls -la gpurun_out/prof_r2a_c3.ncu-rep
