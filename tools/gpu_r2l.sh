#!/bin/bash
# round 2, GPU call L: per-line stall samples of the linking kernel (S09_D6 and the synthetic sequence)
set -x
cd "$GRAFT_REPO_ROOT"
for tag in s09 syn; do
  args=""; [ $tag = syn ] && args="--synthetic 3000"
  timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:link_kernel -s 1 -c 1 -f -o /tmp/link_$tag python tools/link_profile_run.py $args > gpurun_out/r2l_ncu_$tag.log 2>&1
  ls -la /tmp/link_$tag.ncu-rep
  ncu -i /tmp/link_$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2l_link_${tag}_source.csv 2>/dev/null
done
ls -la gpurun_out/r2l*
