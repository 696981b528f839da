# round 2, GPU call F (8 GPUs): bench at N=8, the evidence ladder (one run per GPU), the 4-device handle test,
# the single-process multi-device handle over the whole box
set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n8.log 2> gpurun_out/r2f_bench_n8.err; echo "n8 rc=$?"; tail -c 400 gpurun_out/r2f_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 6 --master-addr 127.0.0.1 --master-port 29513 examples/evidence_ladder.py --kmax 5 --epochs 300 --nlive 200 --cpu-kmax 1 > gpurun_out/r2f_ladder.log 2> gpurun_out/r2f_ladder.err; echo "ladder rc=$?"; cat gpurun_out/r2f_ladder.log
python -m pytest tests/test_gpu_multi.py -m gpu -q -k "multi_device_handle" 2>&1 | tail -2
python tools/multi_handle_rate.py 3 131072 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r2f_bench_n4.log 2> gpurun_out/r2f_bench_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for f in ("r2f_bench_n8","r2f_bench_n4"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], d.get("gather_check",{}).get("pass"), d.get("gather_cost"), (d.get("parity") or {}).get("pass"), d.get("clocks"))
        for k in ("sweep_total_points","latency_ndraw4096","stress"):
            if k in d: print("  ", k, json.dumps(d[k])[:600])
    except Exception as e: print(f, "failed", e)
PY
