#!/bin/bash
tag=${1:-x}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_bh.py -q -m gpu -k "radix or tree_inv or config4" > $out/pytest_sort_$tag.log 2>&1; tail -2 $out/pytest_sort_$tag.log
timeout 300 python tools/bh_timing.py 1048576,16777216 2>&1 | grep -v "theta=0.35" | tail -4
CMD="python bench.py --workload plummer_1m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_bh_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bh_$tag.csv $CMD > $out/ncu_launches_bh_$tag.log 2>&1
echo "ncu launches rc=$?"
$CMD > $out/plain_bh2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bh_walk_group|radix_scatter|radix_hist|tree_split|monopole|gather_bodies" -s 60 -c 20 -f -o $out/prof_bh_$tag $CMD > $out/ncu_full_bh_$tag.log 2>&1
echo "ncu full rc=$?"; tail -2 $out/ncu_full_bh_$tag.log | cut -c1-200
