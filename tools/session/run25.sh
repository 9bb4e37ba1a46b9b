timeout 120 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto and against_oracle" > gpurun_out/pytest_pipe.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_pipe.log
timeout 600 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_pipe2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_pipe2.log
for sp in 0 1; do
KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b25.log 2>&1; tail -1 gpurun_out/b25.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT spare=$sp ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
