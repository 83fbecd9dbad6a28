#!/bin/bash
# round 2, GPU call K: the sixteen-warp linking kernel -- classifier parity tests, timing, section profile (tuning build)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
timeout 600 python tools/link_iter.py > gpurun_out/r2k_link.log 2>&1; cat gpurun_out/r2k_link.log
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so TRI_CLS_PROFILE=1 timeout 600 python tools/link_iter.py --frames 3000 > gpurun_out/r2k_link_prof.log 2>&1; cat gpurun_out/r2k_link_prof.log
