set -x
for lib in "" evidence_b200/alt_tu0.so evidence_b200/alt_td0.so evidence_b200/alt_td1.so; do
  echo "== lib=$lib"
  RVL_LIB=$lib python tools/prof_sweep.py 3 524288 0 | tail -1
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 0 | tail -1
done
