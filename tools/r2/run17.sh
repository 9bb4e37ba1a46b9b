#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fused_ce.py tests/test_gpu_features.py -q -m gpu 2>&1 | tail -40
timeout 300 python tools/r2/nce_tc_time.py 8192 46656 2>&1 | tail -5
timeout 600 python tools/r2/time_modes.py reference reference_ce reference_T0.8p reference_T0.8p_ce reference_T1 reference_T1_ce 2>&1 | tail -8
