#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q -k "lazy or twelve" 2>&1 | tail -3
timeout 600 python tools/lazy_iter.py 2>&1 | tail -9
