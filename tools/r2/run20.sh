#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_logit_kd.py tests/test_gpu_kernel_variants.py tests/test_gpu_forward_mixing.py -q -m gpu -k "kl or log_target or temperature or compute_loss or variants" 2>&1 | tail -15
timeout 600 python tools/r2/time_modes.py kl kl_T0.8 kl_lt_T0.8 2>&1 | tail -4
