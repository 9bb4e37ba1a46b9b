set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for sp in 1 0; do
KD_SPARE_SMS=$sp timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench17_$sp.log 2>&1; tail -1 gpurun_out/bench17_$sp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT spare=$sp ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'stats',d['secondary']['teacher_stats']['kernel_ms'],'per_row',d['secondary']['loca_per_row']['ms'],'kl',d['secondary']['kl_plain_T0.8']['ms'],'loss',d['loss'])"
done
