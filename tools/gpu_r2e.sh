#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
TRI_CLS_PROFILE=1 TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so timeout 600 python tools/cls_debug.py > gpurun_out/r2e_prof.log 2>&1
timeout 600 python tools/bench_classify.py --skip-cpu > gpurun_out/r2e_classify.jsonl 2> gpurun_out/r2e_classify.err
tail -30 gpurun_out/r2e_pytest.log; grep -E "link cycles|stats|mismatch" gpurun_out/r2e_prof.log; cat gpurun_out/r2e_classify.jsonl
