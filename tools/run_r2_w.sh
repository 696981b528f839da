# round 2, GPU call W (8 GPUs): final kernel build at N = 8 (full line), 4 and 2 (headline only), the multi-device handle
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2w_bench_n8.log 2> gpurun_out/r2w_bench_n8.err; echo "n8 rc=$?"; tail -c 300 gpurun_out/r2w_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-parity > gpurun_out/r2w_bench_n4.log 2> gpurun_out/r2w_bench_n4.err; echo "n4 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-parity > gpurun_out/r2w_bench_n2.log 2> gpurun_out/r2w_bench_n2.err; echo "n2 rc=$?"
timeout 300 python tools/multi_handle_rate.py 3 131072 > gpurun_out/r2w_multi_handle.log 2>&1; cat gpurun_out/r2w_multi_handle.log
python - <<'PY'
import json
for f in ("r2w_bench_n2","r2w_bench_n4","r2w_bench_n8"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value %.5g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], d["e2e"]["ms_per_call_percentiles_1_50_99"], d["e2e"]["d2h_bytes_per_step"], d["e2e"]["result_equals_device_path"], d.get("gather_check",{}).get("pass"), d.get("gather_check",{}).get("shared_host_call_equal"), d.get("gather_cost",{}).get("wait_after_kernel_ms_mean_max_over_ranks"), (d.get("parity") or {}).get("pass"), d["roofline"]["frac"])
        for k in ("latency_ndraw4096","stress","sweep_total_points"):
            if k in d: print("  ", k, json.dumps(d[k])[:600])
    except Exception as e: print(f, "failed", e)
PY
