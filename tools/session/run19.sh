PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
run() { timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT $1 ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'stats',d['secondary']['teacher_stats']['kernel_ms'])"; }
KD_SPARE_SMS=2 run queue_nohelper
KD_SPARE_SMS=0 run static
KD_SPARE_SMS=1 run helper
KD_SPARE_SMS=2 run queue_nohelper
