set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-extras > gpurun_out/r2x_bench.log 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2x_bench.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], d["parity"]["pass"], d["parity"]["max_abs"])
PY
