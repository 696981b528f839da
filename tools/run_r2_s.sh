set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
timeout 600 python bench.py --no-extras > gpurun_out/r2s_bench.log 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2s_bench.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], d["parity"]["pass"], d["parity"]["max_abs"])
PY
python tools/prof_sweep.py 2 4096 0 | tail -1
python tools/prof_sweep.py 3 10000 0 | tail -1
