set -x
python -m pytest tests/test_gpu_runner.py -m gpu -q --durations=8 2>&1 | tail -16
