#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
export TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so
{
echo "== DLT f64: 0 default | 6 lazy(2,2) | 7 lazy(3,2) | 8 lazy(2,3) | 9 FPT1 lazy(3,3) | 10 FPT1 lazy(3,4) | 11 FPT1 regs(3,4) | 12 FPT1 lazy (4,5)"
timeout 600 python tools/ab_variants.py --variants 0,6,7,8,9,10,11,12,0
echo "== DLT f32: 0 default (2,3) | 6 lazy(2,3) | 7 lazy(3,3) | 8 lazy(2,4) | 9 lazy(3,4)"
timeout 600 python tools/ab_variants.py --precision f32 --variants 0,6,7,8,9,0
echo "== ray f32: 0 default (3,2) | 4 lazy(3,2) | 5 lazy(3,3) | 6 lazy(2,4) | 7 lazy(4,3)"
timeout 600 python tools/ab_variants.py --mode ray --precision f32 --variants 0,4,5,6,7,0
echo "== ray closed f64: 0 default | 2 lazy(3,2) | 3 lazy(3,3) | 4 FPT1 lazy (3,4)"
timeout 600 python tools/ab_variants.py --mode ray --variants 0,2,3,4,0
echo "== ray LM f64: 0 default | 2 lazy(2,2) | 3 lazy(3,3)"
timeout 600 python tools/ab_variants.py --mode ray --flags 64 --variants 0,2,3
} > gpurun_out/r2g_ab.log 2>&1
grep -v "^+" gpurun_out/r2g_ab.log | sed 's/max |f32.*differ/ .. differ/'
