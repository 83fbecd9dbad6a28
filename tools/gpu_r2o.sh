#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/link_iter.py --frames 100000 2>&1 | awk '!seen[substr($0,1,30)]++'
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
