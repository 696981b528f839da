# round 2, GPU call U (1 GPU): final kernel build -- full GPU tests, smoke, both bench arms, launch list + full ncu
# capture of the bench command, ncu of the latency workload
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench1.log 2> gpurun_out/r2u_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2u_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_bench.csv python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2u_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 4 -c 1 -f -o gpurun_out/prof_r2u_bench_c3 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2u_ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2u_c2 python tools/prof_sweep.py 2 4096 0 > gpurun_out/r2u_ncu2.log 2>&1
ls -la gpurun_out/prof_r2u*
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2u_bench1.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], d["parity"]["pass"])
print(json.dumps(d["latency_ndraw4096"]))
print(json.dumps(d["sweep_total_points"]))
print(json.dumps(d["stress"])[:400])
PY
