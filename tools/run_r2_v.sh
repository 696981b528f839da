set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
  python tools/prof_sweep.py 3 524288 0 | tail -1
  python tools/prof_sweep.py 2 4096 0 | tail -1
done
