# round 2, GPU call Z (1 GPU): the full GPU suite and smoke on the last build
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
