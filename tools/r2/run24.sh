#!/bin/bash
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/round2_bench_n${N}_final.json 2> gpurun_out/round2_bench_n${N}_final.err
tail -3 gpurun_out/round2_bench_n${N}_final.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/round2_bench_n${N}_final.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('parity'))
print(json.dumps(d['secondary'].get('lm_head_kd_vshard'))[:1200])
print(d['e2e'])
PY
