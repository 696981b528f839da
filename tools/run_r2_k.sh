# round 2, GPU call K (1 GPU): final build -- full GPU tests, smoke, bench, launch list + full ncu capture of the
# bench command, ncu of the latency workload
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2k_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench1.log 2> gpurun_out/r2k_bench1.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2k_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 4 -c 1 -f -o gpurun_out/prof_r2k_bench_c3 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2k_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2k_c2 python tools/prof_sweep.py 2 4096 0 > gpurun_out/r2k_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2k_c3_u2 python tools/prof_sweep.py 3 131072 2 > gpurun_out/r2k_ncu3.log 2>&1
ls -la gpurun_out/prof_r2k*
