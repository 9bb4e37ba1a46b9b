timeout 900 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_fused.log 2>&1; tail -15 gpurun_out/pytest_fused.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench22.log 2>&1; tail -1 gpurun_out/bench22.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT ms/step',d['ms_per_step'],'tok/s',d['value'],'emit',d['roofline']['kernel_ms'],d['roofline']['frac'],'stats',d['secondary']['teacher_stats']['kernel_ms'],'per_row',d['secondary']['loca_per_row']['ms'],'kl',d['secondary']['kl_plain_T0.8']['ms'],'clocks',d['clocks'])"
