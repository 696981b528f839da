set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err || exit 1
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 8 -c 1 -f -o gpurun_out/prof_r1c_bench_c2 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu_f.log 2>&1
python tools/prof_sweep.py 3 65536 > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r1c_c3 python tools/prof_sweep.py 3 65536 > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/p3.log
python tools/warp_trace.py 2 4096 > gpurun_out/trace_final.log 2>&1
python tools/warp_trace.py 2 65536 2>/dev/null | head -1 >> gpurun_out/trace_final.log
