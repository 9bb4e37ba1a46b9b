#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/r2/time_gemms.py 2>&1 | tail -5
python tools/lmhead_time.py 2>&1 | tail -5
