#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_tuning.so TRI_CLS_PROFILE=1 timeout 600 python tools/link_iter.py --frames 3000 2>&1 | awk '!seen[substr($0,1,30)]++' > gpurun_out/r2n_link_prof.log; cat gpurun_out/r2n_link_prof.log
timeout 600 python tools/link_iter.py --frames 3000 2>&1 | awk '!seen[substr($0,1,30)]++' > gpurun_out/r2n_link.log; cat gpurun_out/r2n_link.log
timeout 600 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio --clock-control none -k regex:link_kernel -c 1 python tools/link_profile_run.py 2>&1 | grep -E "link_kernel|icc|inst_executed|duration|stalled" > gpurun_out/r2n_icc.log; cat gpurun_out/r2n_icc.log
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -3
