set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for lib in "" evidence_b200/alt_tl0.so "" evidence_b200/alt_tl0.so; do
  echo "== lib=$lib"
  RVL_LIB=$lib python tools/prof_sweep.py 3 524288 0 | tail -1
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 0 | tail -1
done
