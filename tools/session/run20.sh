timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench20.log 2>&1; tail -1 gpurun_out/bench20.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT ms/step',d['ms_per_step'],'tok/s',d['value'],'emit',d['roofline']['kernel_ms'],d['roofline']['frac'],'stats',d['secondary']['teacher_stats']['kernel_ms'],'per_row',d['secondary']['loca_per_row']['ms'],'kl',d['secondary']['kl_plain_T0.8']['ms'],'clocks',d['clocks'])"
