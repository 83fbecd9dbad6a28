# usage: ab_env.sh "<kernel list for profile_kernels --only>" v0 v1 ...   -- TRI_VARIANT A/B of tools/profile_kernels.py
only=$1; shift
for v in "$@"; do echo "TRI_VARIANT=$v"; TRI_VARIANT=$v python tools/profile_kernels.py --frames 100000000 --reps 5 --only $only 2>&1 | sed 's/^/  /'; done
