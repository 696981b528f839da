set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "general_loop or cap_exit or iteration_cap" 2>&1 | tail -30
