#!/bin/bash
# quick linking-kernel iteration: S09_D6 golden check + timings (product build), section profile (tuning build)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/link_iter.py > gpurun_out/r2m_link.log 2>&1; cat gpurun_out/r2m_link.log
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so TRI_CLS_PROFILE=1 timeout 600 python tools/link_iter.py --frames 3000 2>&1 | awk '!seen[substr($0,1,30)]++' > gpurun_out/r2m_link_prof.log; cat gpurun_out/r2m_link_prof.log
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -3
