# round 2, GPU call C (1 GPU): all GPU tests, the pinned-constant loop A/B, the bench, ncu captures
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2c_pytest_gpu.log
python tests/diag/diag_highecc.py 3 2>&1 | tail -7
for lib in "" evidence_b200/variants/librvlnl_nopin.so; do
  tag=${lib:+nopin}; tag=${tag:-pin}
  for args in "2" "2 32" "3" "4"; do
    RVL_LIB=$lib python tools/prof_sweep.py 3 131072 $args 2>&1 | tail -1 | sed "s/^/[$tag] /"
  done
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 2 2>&1 | tail -1 | sed "s/^/[$tag] /"
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 2 32 2>&1 | tail -1 | sed "s/^/[$tag] /"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench1.log 2> gpurun_out/r2c_bench1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2c_bench1.err
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2c_c3 python tools/prof_sweep.py 3 131072 > gpurun_out/r2c_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2c_c2 python tools/prof_sweep.py 2 4096 > gpurun_out/r2c_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
