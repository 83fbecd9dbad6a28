#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python tools/bench_classify.py --skip-cpu 2>&1 | cut -c1-330
timeout 1500 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q 2>&1 | tail -3
