PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
run() { timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT $1 ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'stats',d['secondary']['teacher_stats']['kernel_ms'])"; }
KD_SPARE_SMS=0 run new_spare0
KD_SPARE_SMS=1 run new_spare1
cp $PKG/csrc/logit_kd_cluster.cu /tmp/a.cu; cp $PKG/csrc/logit_kd.cuh /tmp/a.cuh; cp $PKG/csrc/logit_kd.cu /tmp/b.cu; cp include/kd_b200.h /tmp/h.h; cp $PKG/sharded.py /tmp/s.py
cp gpurun_out/oldsrc/logit_kd_cluster.cu $PKG/csrc/; cp gpurun_out/oldsrc/logit_kd.cuh $PKG/csrc/; cp gpurun_out/oldsrc/logit_kd.cu $PKG/csrc/; cp gpurun_out/oldsrc/kd_b200.h include/; cp gpurun_out/oldsrc/sharded.py $PKG/
python $PKG/build.py --force > /dev/null
run old_commit
