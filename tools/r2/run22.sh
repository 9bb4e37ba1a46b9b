#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fused_ce.py tests/test_gpu_module.py -q -m gpu 2>&1 | tail -12
timeout 600 python tools/r2/time_modes.py reference reference_ce reference_T0.8p_ce reference_T1_ce 2>&1 | tail -5
