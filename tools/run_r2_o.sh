set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q -k "fused_gather or bounded" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2o_bench_n2.log 2> gpurun_out/r2o_bench_n2.err; echo "n2 rc=$?"; tail -c 500 gpurun_out/r2o_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --e2e-gather device > gpurun_out/r2o_bench_n2_dev.log 2> gpurun_out/r2o_bench_n2_dev.err; echo "n2 dev rc=$?"
python - <<'PY'
import json
for f in ("r2o_bench_n2","r2o_bench_n2_dev"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], d["e2e"]["ms_per_call_percentiles_1_50_99"], d["e2e"]["d2h_bytes_per_step"], d["e2e"]["result_equals_device_path"], d.get("gather_check"))
    except Exception as e: print(f, "failed", e)
PY
