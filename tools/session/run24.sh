for flags in "" "-DKD_ABLATE_SYNC"; do
KD_EXTRA_NVCC_FLAGS="$flags" python -c "
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import build; build.build(force=True)" > /dev/null 2>&1
KD_SPARE_SMS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b24.log 2>&1; tail -1 gpurun_out/b24.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT flags=$flags ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'])"
done
