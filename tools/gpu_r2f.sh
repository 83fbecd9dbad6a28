#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
TRI_CLS_PROFILE=1 TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so timeout 600 python tools/cls_debug.py > gpurun_out/r2f_prof.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
tail -5 gpurun_out/r2f_pytest.log; grep -E "link cycles|stats|mismatch" gpurun_out/r2f_prof.log; tail -3 gpurun_out/r2f_bench.err; head -c 600 gpurun_out/r2f_bench_ref.json
