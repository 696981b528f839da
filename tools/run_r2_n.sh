set -x
for lib in "" evidence_b200/variants/librvlnl_nopp.so; do
  tag=${lib:+nopp}; tag=${tag:-pp}
  for args in "2" "4"; do
    RVL_LIB=$lib python tools/prof_sweep.py 3 131072 $args 2>&1 | tail -1 | sed "s/^/[$tag] /"
  done
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 2 2>&1 | tail -1 | sed "s/^/[$tag] /"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
python tests/diag/diag_highecc.py 3 2>&1 | tail -5 | cut -c1-200
