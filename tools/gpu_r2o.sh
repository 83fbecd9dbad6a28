#!/bin/bash
cd "$GRAFT_REPO_ROOT"
echo "== global-memory sort forced (cap 256)"
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_G.so timeout 900 python -m pytest tests/test_gpu_classify.py tests/test_gpu_ref.py -m gpu -x -q 2>&1 | tail -3
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_G.so timeout 600 python tools/link_iter.py --frames 20000 2>&1 | awk '!seen[substr($0,1,30)]++'
echo "== product"
timeout 600 python tools/link_iter.py --frames 100000 2>&1 | awk '!seen[substr($0,1,30)]++'
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -3
