#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_features.py -q -m gpu -x 2>&1 | tail -25
timeout 300 python tools/r2/nce_tc_time.py 8192 46656 2>&1 | tail -5
