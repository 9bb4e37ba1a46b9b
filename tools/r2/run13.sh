#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/r2/nce_kernel_time.py 2>&1 | tail -2
KD_NCE_FUSED=0 python tools/r2/nce_kernel_time.py 2>&1 | tail -2
python tools/nce_times.py 2>&1 | tail -1
