#!/bin/bash
# ncu --set full of one launch of the walk kernel (after the plain run exited 0).  usage: gpu_ncu_walk.sh TAG GROUP [KERNEL_REGEX]
tag=${1:-r2h}; gs=${2:-64}; kern=${3:-bh_walk}
out=gpurun_out; mkdir -p $out
timeout 120 python tools/walk_probe.py 1048576 $gs > $out/walk_probe_$tag.log 2>&1 || { echo "plain run failed"; cat $out/walk_probe_$tag.log; exit 1; }
cat $out/walk_probe_$tag.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kern -s 3 -c 1 -f -o $out/prof_walk_$tag python tools/walk_probe.py 1048576 $gs > $out/ncu_walk_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 $out/ncu_walk_$tag.log
