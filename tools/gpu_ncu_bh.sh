#!/bin/bash
tag=${1:-x}
out=gpurun_out; mkdir -p $out
CMD="python bench.py --workload plummer_1m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_bh2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bh_walk_group|radix_scatter|tree_split|monopole" -s 128 -c 32 -f -o $out/prof_bh_$tag $CMD > $out/ncu_full_bh_$tag.log 2>&1
echo "ncu full rc=$?"; tail -3 $out/ncu_full_bh_$tag.log
