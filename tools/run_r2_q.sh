set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.log 2> gpurun_out/r2q_bench_n2.err; echo "n2 rc=$?"; tail -c 400 gpurun_out/r2q_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2q_bench_n2.log") if l.startswith("{")][-1])
print("value %.5g"%d["value"], "e2e %.5g"%d["e2e"]["value"], d["gather_check"], d["parity"]["pass"])
print(json.dumps(d["latency_ndraw4096"]))
print(json.dumps(d["stress"])[:300])
PY
