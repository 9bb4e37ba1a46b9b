timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/b40.log 2>&1; tail -1 gpurun_out/b40.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT ms/step %.3f tok/s %.0f emit %.3f frac %.3f stats %.3f (%.3f) whole %.3f per_row %.3f kl %.3f nce %.0f mse %.3f loss %.9f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['secondary']['teacher_stats']['kernel_ms'],d['secondary']['teacher_stats']['frac'],d['secondary']['whole_step']['frac'],d['secondary']['loca_per_row']['ms'],d['secondary']['kl_plain_T0.8']['ms'],d['secondary']['info_nce_64x729x1152']['us'],d['secondary']['feature_mse_65536x3584']['ms'],d['loss']))"
done
