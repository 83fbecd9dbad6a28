#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for lib in E5 E6 E8; do
echo "== $lib"; TRI_B200_LIB=$PWD/3d-reconstruction-triangulation_b200/libtri_b200_$lib.so timeout 600 python tools/link_iter.py --frames 100000 2>&1 | awk '!seen[substr($0,1,30)]++' | grep -v golden | tail -3
done
echo "== current"; timeout 600 python tools/link_iter.py --frames 100000 2>&1 | awk '!seen[substr($0,1,30)]++' | grep -v golden | tail -3
