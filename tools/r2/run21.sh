#!/bin/bash
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
SOAK_ONLY=ce timeout 900 python tools/soak.py 400 2>&1 | tail -12
SOAK_ONLY=log_target timeout 600 python tools/soak.py 400 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_logit_kd.py -q -m gpu -k "near_one_hot" 2>&1 | tail -4
