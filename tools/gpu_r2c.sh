#!/bin/bash
# round 2, GPU call C: the single-warp linking kernel -- parity (oracle, goldens, _ref) and timing
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 600 python tools/bench_classify.py > gpurun_out/r2c_classify.jsonl 2> gpurun_out/r2c_classify.err
tail -25 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_classify.jsonl; tail -5 gpurun_out/r2c_classify.err
