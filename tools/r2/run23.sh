#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py > gpurun_out/round2_bench_n1_final.json 2> gpurun_out/round2_bench_n1_final.err; tail -2 gpurun_out/round2_bench_n1_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/round2_bench_reference_arm.json 2>/dev/null
timeout 600 python bench.py --mode per_row --no-secondary --no-cpu-baseline > gpurun_out/round2_bench_n1_per_row.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/round2_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e --no-parity > gpurun_out/round2_launches_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/round2_nce_tc_launches.csv python tools/profile_ops.py nce_tc > gpurun_out/round2_nce_tc_launches_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/round2_bench_n1_final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('parity',{}).get('ok'), d['clocks'])
for k,v in d['secondary'].items(): print(k, json.dumps(v)[:260])
print(d['cpu_baseline']); print(d['e2e'])
PY
