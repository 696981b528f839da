# round 2, GPU call G (1 GPU): full GPU tests after the sampler / post-processing changes, post-processing bench,
# work-list sweep on the latency workload
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2g_pytest_gpu.log
python tools/post_bench.py > gpurun_out/r2g_post_bench.log 2>&1; cat gpurun_out/r2g_post_bench.log
python tools/sched_sweep.py 2 4096 200 > gpurun_out/r2g_sched_sweep.log 2>&1; cat gpurun_out/r2g_sched_sweep.log
python tools/sampler_rate.py > gpurun_out/r2g_sampler_rate.log 2>&1; tail -5 gpurun_out/r2g_sampler_rate.log
