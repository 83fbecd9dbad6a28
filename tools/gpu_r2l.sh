#!/bin/bash
# per-instruction stall samples of the linking kernel (S09_D6)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:link_kernel -s 1 -c 1 -f -o /tmp/link_s09 python tools/link_profile_run.py > gpurun_out/r2l_ncu_s09.log 2>&1
ncu -i /tmp/link_s09.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2l_link_s09_source.csv 2>/dev/null
ls -la gpurun_out/r2l*
