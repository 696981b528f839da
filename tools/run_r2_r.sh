set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2r_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_ref.log 2> gpurun_out/r2r_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2r_ref.log
timeout 900 python bench.py > gpurun_out/r2r_bench.log 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2r_bench.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "e2e %.5g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], d["parity"]["pass"], d["clocks"])
print(json.dumps(d.get("latency_ndraw4096")))
PY
