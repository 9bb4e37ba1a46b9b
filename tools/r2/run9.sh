#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_lm_head.py -x -q -m gpu 2>&1 | tail -15
