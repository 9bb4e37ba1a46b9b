#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_logit_kd.py -x -q -m gpu 2>&1 | tail -5
echo "== default"; python tools/r2/time_modes.py per_row kl per_row_T0.8p
echo "== trace"; KD_TRACE_DUMP=$PWD/gpurun_out/trace_pr.bin KD_B200_LIB=$PWD/tools/variants/trace/libkd_b200.so ROWS=8192 python tools/r2/time_modes.py per_row
python tools/trace_fused.py gpurun_out/trace_pr.bin; rm -f gpurun_out/trace_pr.bin
