for rs in 10 30 60 100; do
KD_HELPER_RESERVE=$rs timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b41.log 2>&1; tail -1 gpurun_out/b41.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT reserve=$rs ms/step %.3f tok/s %.0f emit %.3f frac %.3f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
done
KD_SPARE_SMS=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b41.log 2>&1; tail -1 gpurun_out/b41.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT nohelper ms/step %.3f tok/s %.0f emit %.3f frac %.3f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
