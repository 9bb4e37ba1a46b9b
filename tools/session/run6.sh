set -x
timeout 1200 python -m pytest tests/test_gpu_logit_kd.py -m gpu -q -k "not other_dtypes and not grad_output and not full_vocab" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for cfg in 0; do
KD_CLUSTER_CONFIG=$cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench6_$cfg.log 2>&1; tail -1 gpurun_out/bench6_$cfg.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'tok/s',d['value']); print('roof',d['roofline']['kernel_ms'],d['roofline']['frac']); 
for k,v in d['secondary'].items(): print(k,v)"
done
python tools/profile_emit.py 4096 reference 2 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:row_emit_cluster -s 1 -c 1 -o gpurun_out/emit_r4 python tools/profile_emit.py 4096 reference 2 > gpurun_out/ncu_emit.log 2>&1
tail -2 gpurun_out/ncu_emit.log
