timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for tm in 0 1; do
KD_STATS_TMA=$tm timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/b34_$tm.log 2>&1; tail -1 gpurun_out/b34_$tm.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT tma=$tm ms/step %.3f tok/s %.0f emit %.3f frac %.3f stats %.3f (%.3f) per_row %.3f kl %.3f loss %.9f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['secondary']['teacher_stats']['kernel_ms'],d['secondary']['teacher_stats']['frac'],d['secondary']['loca_per_row']['ms'],d['secondary']['kl_plain_T0.8']['ms'],d['loss']))"
done
