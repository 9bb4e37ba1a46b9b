timeout 600 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -x -k "auto" > gpurun_out/pytest_pipe2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_pipe2.log
for sp in 0 1; do
KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b26.log 2>&1; tail -1 gpurun_out/b26.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT spare=$sp ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
python tools/profile_emit.py 4096 reference 2 > gpurun_out/plain.log 2>&1 && \
KD_SPARE_SMS=0 ncu --set full --clock-control none --import-source on -k regex:row_emit_pipelined -s 1 -c 1 -o gpurun_out/pipe_r1 python tools/profile_emit.py 4096 reference 2 > gpurun_out/ncu_pipe.log 2>&1
tail -2 gpurun_out/ncu_pipe.log
