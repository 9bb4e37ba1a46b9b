for nc in "" 1; do
KD_BENCH_NO_CLOCKS=$nc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/bench_g2.log 2>&1; grep -E "^\{" gpurun_out/bench_g2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('noclocks=$nc ms/step %.3f value %.0f clocks %s'%(d['ms_per_step'], d['value'], d['clocks']))"
done
