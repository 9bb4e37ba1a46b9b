#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_features.py -q -m gpu 2>&1 | tail -25
timeout 300 python tools/r2/nce_tc_time.py 8192 46656 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/nce_tc_launches.csv python tools/r2/nce_tc_time.py 46656 > gpurun_out/nce_tc_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/nce_tc_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:][-14:]: print(r[ki][:90], r[vi])
PY
