set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2i_pytest_multi.log 2>&1; tail -3 gpurun_out/r2i_pytest_multi.log
