#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_module.py tests/test_gpu_features.py -q -m gpu 2>&1 | tail -30
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; tail -3 gpurun_out/r2_bench18.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench18.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d.get('parity',{}).get('ok'))
for k,v in d['secondary'].items(): print(k, json.dumps(v)[:300])
PY
