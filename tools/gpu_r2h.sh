#!/bin/bash
# round 2, GPU call H: ray FP32 single-pass tile variants, parity of the changed tiles, then ncu: full captures of the
# shipped kernels and the launch list of a short bench run (each after the plain command has exited 0)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
{
echo "== ray f32 (single pass): 0 default (3,2) | 4 lazy(3,2) | 5 lazy(3,3) | 6 lazy(2,4) | 7 lazy(4,3)"
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so timeout 600 python tools/ab_variants.py --mode ray --precision f32 --variants 0,4,5,6,7,0
echo "== DLT f32 default (product library)"
timeout 600 python tools/ab_variants.py --precision f32 --variants 0
} > gpurun_out/r2h_ab.log 2>&1
timeout 600 python tools/profile_kernels.py --frames 100000000 --reps 3 --only dlt_f64,dlt_f32,ray_closed_f64,ray_f32,ray_f64 > gpurun_out/r2h_kernels.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -c 10 -f -o gpurun_out/prof_r2_batch python tools/profile_kernels.py --frames 100000000 --reps 1 --only dlt_f64,dlt_f32,ray_closed_f64,ray_f32,ray_f64 > gpurun_out/r2h_ncu_batch.log 2>&1
timeout 300 python tools/link_profile_run.py > gpurun_out/r2h_link.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"link_kernel|enumerate_kernel" -c 4 -f -o gpurun_out/prof_r2_classify python tools/link_profile_run.py > gpurun_out/r2h_ncu_cls.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2h_bench_short.json 2> gpurun_out/r2h_bench_short.err &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2h_ncu_launch.log 2>&1
tail -3 gpurun_out/r2h_pytest.log; grep -v "^+" gpurun_out/r2h_ab.log | sed 's/max |f32.*differ/ /'; cat gpurun_out/r2h_kernels.log; ls -la gpurun_out/*.ncu-rep | tail -3
