for v in ab0_c0 ab2_c0 ab2_c1 ab1_c1 ab3_c1 ab2_c2 ab4_c2; do for f in 1 2; do
KD_B200_LIB=$PWD/tools/variants/libkd_$v.so KD_FUSED=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b28.log 2>&1; tail -1 gpurun_out/b28.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT $v fused=$f ms/step %.3f emit %.3f frac %.3f loss %.9f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['loss']))"
done; done
timeout 600 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_poly.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_poly.log
KD_FUSED=2 timeout 600 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_poly2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_poly2.log
