set -x
PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
for flags in "-DKD_ABLATE_MUFU" "-DKD_ABLATE_LDS" "-DKD_ABLATE_MUFU -DKD_ABLATE_LDS"; do
KD_EXTRA_NVCC_FLAGS="$flags" python $PKG/build.py --force > /dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/bench14.log 2>&1; tail -1 gpurun_out/bench14.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT $flags ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'stats',d['secondary']['teacher_stats']['kernel_ms'])"
done
