#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/r2/dbg_big.py reference 2>&1 | tail -8
python tools/r2/dbg_big.py per_row 2>&1 | tail -8
