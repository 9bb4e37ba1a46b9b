#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_logit_kd.py -q -m gpu --tb=short -k "benchmark_shape" 2>&1 | grep -v "Warning" | grep -E "^E|^tests|passed|failed" | head -30
