N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for g in ${2:-fused nccl}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 10 --no-extras --gather $g 2> gpurun_out/bm_${N}_$g.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'])"
done
