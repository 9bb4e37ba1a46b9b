set -x
PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
for cfg in "1024 1" "1024 2" "768 2" "512 2"; do
set -- $cfg
KD_EXTRA_NVCC_FLAGS="-DKD_CLUSTER_THREADS=$1 -DKD_SWEEP_UNROLL=$2" python $PKG/build.py --force > /dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench13.log 2>&1; tail -1 gpurun_out/bench13.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT threads $1 unroll $2 ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'per_row',d['secondary']['loca_per_row']['ms'],'kl',d['secondary']['kl_plain_T0.8']['ms'])"
done
