#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:enumerate_kernel -s 1 -c 1 -f -o /tmp/enum_syn python tools/link_profile_run.py --synthetic 20000 > gpurun_out/r2q_ncu.log 2>&1
ncu -i /tmp/enum_syn.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2q_enum_source.csv 2>/dev/null
ncu -i /tmp/enum_syn.ncu-rep --page details > gpurun_out/r2q_enum_details.txt 2>/dev/null
