set -x
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_g$N.log 2>&1; tail -2 gpurun_out/bench_g$N.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --mode per_row > gpurun_out/bench_g${N}_row.log 2>&1; tail -1 gpurun_out/bench_g${N}_row.log | cut -c1-600
