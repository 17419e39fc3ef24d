#!/bin/bash
# Round 2, call J (1 GPU): launch lists (duration + DRAM bytes) of the two 1-GPU bench workloads and ncu --set full of the
# two dominant kernels (each after the plain command exited 0).
tag=${1:-r2j}
out=gpurun_out; mkdir -p $out
D="python bench.py --workload plummer_1m_direct --steps 2 --warmup 3 --no-cpu-baseline --no-bh --e2e-steps 1"
B="python bench.py --workload plummer_1m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 $D > $out/plain_direct_$tag.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file $out/launches_direct_$tag.csv $D > $out/ncu_launches_direct_$tag.log 2>&1
echo "direct launches rc=$?"
timeout 300 $B > $out/plain_bh_$tag.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file $out/launches_bh_$tag.csv $B > $out/ncu_launches_bh_$tag.log 2>&1
echo "bh launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:direct_packed -s 2 -c 1 -f -o $out/prof_direct_$tag $D > $out/ncu_full_direct_$tag.log 2>&1
echo "direct full rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bh_walk_group -s 3 -c 1 -f -o $out/prof_walk_$tag $B > $out/ncu_full_walk_$tag.log 2>&1
echo "walk full rc=$?"
ls -la $out/*$tag*
