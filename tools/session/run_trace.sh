set -x
PKG=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
KD_EXTRA_NVCC_FLAGS="-DKD_TRACE -DKD_CLUSTER_THREADS=1024" python $PKG/build.py --force > /dev/null
KD_TRACE_DUMP=gpurun_out/trace.bin python tools/profile_emit.py 16384 reference 4 > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python tools/trace_emit.py gpurun_out/trace.bin 32
