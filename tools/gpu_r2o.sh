#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/link_iter.py --frames 20000 2>&1 | awk '!seen[substr($0,1,30)]++'
timeout 1500 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q 2>&1 | tail -3
