timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cs in 8 9; do for sp in 0 1; do
KD_CLUSTER_SIZE=$cs KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b37.log 2>&1; tail -1 gpurun_out/b37.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT cs=$cs spare=$sp ms/step %.3f tok/s %.0f emit %.3f frac %.3f loss %.9f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['loss']))"
done; done
