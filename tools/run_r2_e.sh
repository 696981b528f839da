# round 2, GPU call E (2 GPUs): multi-GPU tests (fused gather device + host calls, rvl_create_multi, ladder),
# bench at N=2 (both gathers), bench at N=1 on the same box for the efficiency
set -x
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2e_pytest_multi.log 2>&1; tail -5 gpurun_out/r2e_pytest_multi.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2e_bench_n1.log 2> gpurun_out/r2e_bench_n1.err; echo "n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.log 2> gpurun_out/r2e_bench_n2.err; echo "n2 rc=$?"; tail -c 600 gpurun_out/r2e_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --gather nccl > gpurun_out/r2e_bench_n2_nccl.log 2> gpurun_out/r2e_bench_n2_nccl.err; echo "n2 nccl rc=$?"
python - <<'PY'
import json
for f in ("r2e_bench_n1","r2e_bench_n2","r2e_bench_n2_nccl"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], d.get("gather_check"), d.get("gather_cost"), (d.get("parity") or {}).get("pass"))
    except Exception as e: print(f, "failed", e)
PY
