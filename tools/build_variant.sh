#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...> : libkd_b200.so with logit_kd_fused.cu recompiled under extra flags,
# written to tools/variants/<name>/libkd_b200.so (select it with KD_B200_LIB=...).  Tuning / tracing experiments only.
set -e
name=$1; shift
pkg=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
out=tools/variants/$name
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $pkg/csrc/logit_kd_fused.cu -o $out/logit_kd_fused.o
objs=$(ls $pkg/build/*.o | grep -v logit_kd_fused.o)
nvcc -shared -o $out/libkd_b200.so $objs $out/logit_kd_fused.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo $out/libkd_b200.so
