#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/r2/dbg_pr.py 2>&1 | tail -30
echo "== default"; python tools/r2/time_modes.py per_row kl per_row_T0.8p
