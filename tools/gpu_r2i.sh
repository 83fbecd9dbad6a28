#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 2400 python -m pytest tests/test_gpu_classify.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log
