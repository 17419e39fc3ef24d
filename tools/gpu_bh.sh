#!/bin/bash
# BH on the GPU: parity tests, optional timing sweep, BH bench line + ncu launch list + full capture of the walk.
tag=${1:-bh}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_bh.py -q -m gpu --timeout 600 > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 $out/pytest_bh_$tag.log
[ "$2" = "timing" ] && { timeout 600 python tools/bh_timing.py > $out/bh_timing_$tag.log 2>&1; tail -30 $out/bh_timing_$tag.log; }
if [ "$2" = "bench" ]; then
  timeout 600 python bench.py --workload plummer_1m_bh --steps 20 --warmup 5 > $out/bench_bh_$tag.json 2> $out/bench_bh_$tag.err; echo "bench rc=$?"; cat $out/bench_bh_$tag.json; tail -3 $out/bench_bh_$tag.err
  CMD="python bench.py --workload plummer_1m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
  $CMD > $out/plain_bh_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_bh_$tag.csv $CMD > $out/ncu_launches_bh_$tag.log 2>&1
  echo "ncu launches rc=$?"
  $CMD > $out/plain_bh2_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"bh_walk_group|radix_scatter|tree_split|monopole|gather_bodies" -s 400 -c 40 -f -o $out/prof_bh_$tag $CMD > $out/ncu_full_bh_$tag.log 2>&1
  echo "ncu full rc=$?"
fi
