timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --force-streaming > gpurun_out/bench16.log 2>&1; tail -1 gpurun_out/bench16.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT streaming emit',d['roofline']['kernel_ms'])"
