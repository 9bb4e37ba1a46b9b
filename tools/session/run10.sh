set -x
PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for u in 2 3 4; do
KD_EXTRA_NVCC_FLAGS="-DKD_SWEEP_UNROLL=$u" python $PKG/build.py --force > /dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench10_$u.log 2>&1; tail -1 gpurun_out/bench10_$u.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('UNROLL $u ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'per_row',d['secondary']['loca_per_row']['ms'],'kl',d['secondary']['kl_plain_T0.8']['ms'])"
done
