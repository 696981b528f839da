# round 2, GPU call Y (1 GPU): last build -- bench, launch list and full ncu capture of the bench command
set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench1.log 2> gpurun_out/r2y_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2y_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches_bench.csv python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2y_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 4 -c 1 -f -o gpurun_out/prof_r2y_bench_c3 python bench.py --steps 3 --warmup 3 --batch 524288 --no-extras --no-parity > gpurun_out/r2y_ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rv_lnl -s 3 -c 1 -f -o gpurun_out/prof_r2y_c2 python tools/prof_sweep.py 2 4096 0 > gpurun_out/r2y_ncu2.log 2>&1
ls -la gpurun_out/prof_r2y*
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2y_bench1.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "frac", d["roofline"]["frac"], d["parity"]["pass"])
print(json.dumps(d["latency_ndraw4096"]))
print(json.dumps(d["sweep_total_points"]))
print(json.dumps(d["stress"])[:400])
PY
