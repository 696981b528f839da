# round 2, GPU call P (8 GPUs): final build at N = 8 and 4 (shared-host e2e), N = 8 with the device gather for comparison
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n8.log 2> gpurun_out/r2p_bench_n8.err; echo "n8 rc=$?"; tail -c 300 gpurun_out/r2p_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-parity --e2e-gather device > gpurun_out/r2p_bench_n8_dev.log 2> gpurun_out/r2p_bench_n8_dev.err; echo "n8 dev rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-parity > gpurun_out/r2p_bench_n4.log 2> gpurun_out/r2p_bench_n4.err; echo "n4 rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-parity > gpurun_out/r2p_bench_n1.log 2> gpurun_out/r2p_bench_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for f in ("r2p_bench_n1","r2p_bench_n4","r2p_bench_n8","r2p_bench_n8_dev"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value %.5g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], d["e2e"]["ms_per_call_percentiles_1_50_99"], d["e2e"]["d2h_bytes_per_step"], d["e2e"]["result_equals_device_path"], d.get("gather_check",{}).get("pass"), d.get("gather_check",{}).get("shared_host_call_equal"), d.get("gather_cost",{}).get("wait_after_kernel_ms_mean_max_over_ranks"), (d.get("parity") or {}).get("pass"))
        for k in ("latency_ndraw4096","stress"):
            if k in d: print("  ", k, json.dumps(d[k])[:500])
    except Exception as e: print(f, "failed", e)
PY
