#!/bin/bash
cd $GRAFT_REPO_ROOT
echo "== nocand"; KD_B200_LIB=$PWD/tools/variants/nocand/libkd_b200.so python tools/r2/time_modes.py per_row
echo "== trace"; KD_TRACE_DUMP=$PWD/gpurun_out/trace_pr.bin KD_B200_LIB=$PWD/tools/variants/trace/libkd_b200.so ROWS=8192 python tools/r2/time_modes.py per_row
python tools/trace_fused.py gpurun_out/trace_pr.bin; rm -f gpurun_out/trace_pr.bin
echo "== trace kl"; KD_TRACE_DUMP=$PWD/gpurun_out/trace_pr.bin KD_B200_LIB=$PWD/tools/variants/trace/libkd_b200.so ROWS=8192 python tools/r2/time_modes.py kl
python tools/trace_fused.py gpurun_out/trace_pr.bin; rm -f gpurun_out/trace_pr.bin
