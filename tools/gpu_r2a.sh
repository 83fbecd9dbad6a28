#!/bin/bash
# round 2, GPU call A: parity after the tile-API refactor, then the FP64 / FP32 DLT variants and a first bench line
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
export TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so
timeout 600 python tools/ab_variants.py --variants 0,1,2,3,4,5,6,0 > gpurun_out/r2a_ab_f64.log 2>&1
timeout 600 python tools/ab_variants.py --precision f32 --variants 0,1,2,3,4,5,0 > gpurun_out/r2a_ab_f32.log 2>&1
unset TRI_B200_LIB
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_ab_f64.log gpurun_out/r2a_ab_f32.log; head -c 1500 gpurun_out/r2a_bench.json
