set -x
python -m pytest tests/test_fip.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
python tools/post_bench.py 2>&1 | tail -8
python tests/diag/fip_bench.py 200000 3 2>&1 | tail -6
