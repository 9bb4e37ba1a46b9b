for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1; done
for i in 1 2 3 4 5 6; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RUN $i ms/step %.3f tok/s %.0f emit %.3f frac %.3f sm_mhz %s'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['clocks']['sm_mhz']))"
done
