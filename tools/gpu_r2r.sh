#!/bin/bash
# round 2, GPU call R: full GPU test suite, the bench line, the reference arm, then ncu (full captures exported to CSV on the box --
# the .ncu-rep files are too large to travel back -- and the launch list of a short bench run, this library's kernels only)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err; echo "reference rc=$?"
timeout 600 python tools/profile_kernels.py --frames 100000000 --reps 3 --only dlt_f64,dlt_f32,ray_closed_f64,ray_f32,ray_f64 > gpurun_out/r2r_kernels.log 2>&1
timeout 600 python tools/link_iter.py --frames 100000 > gpurun_out/r2r_link.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -c 10 -f -o /tmp/prof_r2_batch python tools/profile_kernels.py --frames 100000000 --reps 1 --only dlt_f64,dlt_f32,ray_closed_f64,ray_f32,ray_f64 > gpurun_out/r2r_ncu_batch.log 2>&1
ncu -i /tmp/prof_r2_batch.ncu-rep --page raw --csv > gpurun_out/r2_ncu_batch_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"link_kernel|enumerate_kernel" -c 4 -f -o /tmp/prof_r2_classify python tools/link_profile_run.py > gpurun_out/r2r_ncu_cls.log 2>&1
ncu -i /tmp/prof_r2_classify.ncu-rep --page raw --csv > gpurun_out/r2_ncu_classify_raw.csv 2>/dev/null
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stream_kernel|batch_kernel|batch_single_kernel|chunk_kernel|enumerate_kernel|link_kernel|ray_reference_kernel|subsets_kernel|dist_from_ray_kernel" -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2r_ncu_launch.log 2>&1
ls -la gpurun_out | tail -14; tail -3 gpurun_out/r2r_pytest.log; cat gpurun_out/r2r_kernels.log gpurun_out/r2r_link.log; tail -2 gpurun_out/r2r_bench.err
