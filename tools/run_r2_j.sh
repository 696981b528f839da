set -x
for lib in "" evidence_b200/variants/librvlnl_norot.so; do
  tag=${lib:+norot}; tag=${tag:-rot}
  for args in "2" "4"; do
    RVL_LIB=$lib python tools/prof_sweep.py 3 131072 $args 2>&1 | tail -1 | sed "s/^/[$tag] /"
  done
  RVL_LIB=$lib python tools/prof_sweep.py 2 4096 2 2>&1 | tail -1 | sed "s/^/[$tag] /"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench1.log 2> gpurun_out/r2j_bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2j_bench1.log") if l.startswith("{")][-1])
print("value %.4g"%d["value"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
p=d["parity"]; print("parity", p["pass"], p["max_abs"], {k:(v["max_abs"],v["bit_identical"]) for k,v in p["sets"].items()})
print(d["latency_ndraw4096"]["kernel_us"], d["latency_ndraw4096"]["frac_of_fp64_peak"], d["stress"]["frac_of_fp64_peak"])
PY
