#!/bin/bash
# round 2, GPU call B: TMA feeder vs cp.async feeder, masking forms; parity of the new default
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
export TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so
{
echo "== DLT f64: 0 hi-select stream(2,2) | 1 TMA(2,2) | 2 TMA(3,2) | 3 table | 4 select | 5 hi-select + smem addends"
timeout 600 python tools/ab_variants.py --variants 0,1,2,3,4,5,0
echo "== DLT f32: 0 stream(2,3) | 1 TMA(2,3) | 2 TMA(3,3) | 3 TMA(3,2) | 4 TMA(4,2) | 5 TMA(3,3)+smem consts"
timeout 600 python tools/ab_variants.py --precision f32 --variants 0,1,2,3,4,5,0
echo "== ray f32: 0 stream(3,2) | 1 TMA(3,2) | 2 TMA(4,2) | 3 TMA(3,3)"
timeout 600 python tools/ab_variants.py --mode ray --precision f32 --variants 0,1,2,3,0
echo "== ray closed f64: 0 stream | 1 TMA"
timeout 600 python tools/ab_variants.py --mode ray --flags 8 --variants 0,1,0
echo "== ray LM f64: 0 stream | 1 TMA"
timeout 600 python tools/ab_variants.py --mode ray --variants 0,1
echo "== stream probe: 0 stream(3,2) | 1 TMA(3,2) | 2 TMA(4,2)"
timeout 600 python tools/ab_variants.py --precision f32 --flags 1073741824 --variants 0,1,2,0
} > gpurun_out/r2b_ab.log 2>&1
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_ab.log
