run() { N=$1; g=$2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 10 --no-extras --gather $g 2> gpurun_out/bs_${N}_$g.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'])"; }
run 8 fused
run 8 nccl
run 4 fused
run 2 fused
