# round 2, GPU call D (1 GPU): full GPU test suite, sampler throughput, bench (both arms), launch list +
# full ncu capture of the bench command, sanitizer passes
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2d_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/sampler_rate.py > gpurun_out/r2d_sampler_rate.log 2>&1; cat gpurun_out/r2d_sampler_rate.log | tail -6
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench1.log 2> gpurun_out/r2d_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2d_bench1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_ref.log 2> gpurun_out/r2d_ref.err; echo "ref rc=$?"
python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2d_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 4 -c 1 -f -o gpurun_out/prof_r2d_bench_c3 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2d_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2d_c2 python tools/prof_sweep.py 2 4096 0 > gpurun_out/r2d_ncu2.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py > gpurun_out/r2d_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2d_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > gpurun_out/r2d_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2d_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_probe.py > gpurun_out/r2d_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2d_synccheck.log
ls -la gpurun_out/*.ncu-rep | tail -3
