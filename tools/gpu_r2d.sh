#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python tools/cls_debug.py > gpurun_out/r2d_debug.log 2>&1
TRI_CLS_PROFILE=1 TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so timeout 600 python tools/cls_debug.py --frames 3000 > gpurun_out/r2d_prof.log 2>&1
cat gpurun_out/r2d_debug.log; grep -E "link cycles|stats" gpurun_out/r2d_prof.log
