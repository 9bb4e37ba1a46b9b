timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for sp in 0 1; do
KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b38.log 2>&1; tail -1 gpurun_out/b38.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT spare=$sp ms/step %.3f tok/s %.0f emit %.3f frac %.3f loss %.9f'%(d['ms_per_step'],d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['loss']))"
done
KD_SPARE_SMS=0 KD_B200_LIB=$PWD/tools/variants/libkd_trace.so KD_TRACE_DUMP=gpurun_out/trace_fused.bin timeout 120 python tools/profile_emit.py 16384 reference 4 > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python tools/trace_fused.py gpurun_out/trace_fused.bin 2>&1 | head -24
