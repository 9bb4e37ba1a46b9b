N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_g$N.log 2>&1; echo rc=$?; grep -E "^\{|\[bench\]" gpurun_out/bench_g$N.log | cut -c1-900
