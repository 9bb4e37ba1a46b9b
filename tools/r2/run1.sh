#!/bin/bash
# first GPU pass of round 2: per-row single-pass parity + timing
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_logit_kd.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py --mode per_row --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -3
