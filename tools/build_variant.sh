#!/bin/bash
# tools/build_variant.sh <name> [--src file.cu] <extra nvcc flags...> : libkd_b200.so with ONE source (default logit_kd_fused.cu)
# recompiled under extra flags, written to tools/variants/<name>/libkd_b200.so (select it with KD_B200_LIB=...).
# Tuning / tracing experiments only; the other objects come from the package's build/ directory (build the package first).
set -e
name=$1; shift
src=logit_kd_fused.cu
if [ "$1" = "--src" ]; then src=$2; shift 2; fi
pkg=knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200
out=tools/variants/$name
obj=${src%.cu}.o
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $pkg/csrc/$src -o $out/$obj
objs=$(ls $pkg/build/*.o | grep -v "/$obj")
nvcc -shared -o $out/libkd_b200.so $objs $out/$obj -gencode arch=compute_100a,code=sm_100a -lcudart
echo $out/libkd_b200.so
