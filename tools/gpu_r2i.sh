#!/bin/bash
tag=${1:-r2i}
out=gpurun_out; mkdir -p $out
for split in 0 1 2; do
  echo "NBODY_WALK=1 NBODY_WS_SPLIT=$split" >> $out/bh_walk_ab_$tag.log
  NBODY_WALK=1 NBODY_WS_SPLIT=$split timeout 300 python tools/bh_timing.py 1048576 walk 2>&1 | grep "theta=0.25" >> $out/bh_walk_ab_$tag.log
done
cat $out/bh_walk_ab_$tag.log
