#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -15
