timeout 600 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_pp.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_pp.log
for te in 0 2 4 6; do for sp in 1; do
KD_TOKEN_EARLY=$te KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b32.log 2>&1; tail -1 gpurun_out/b32.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT early=$te spare=$sp ms/step %.3f emit %.3f frac %.3f loss %.9f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['loss']))"
done; done
KD_SPARE_SMS=0 KD_B200_LIB=$PWD/tools/variants/libkd_trace.so KD_TRACE_DUMP=gpurun_out/trace_fused.bin timeout 120 python tools/profile_emit.py 16384 reference 4 > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python tools/trace_fused.py gpurun_out/trace_fused.bin | head -22
