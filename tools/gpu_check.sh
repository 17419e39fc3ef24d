#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the default bench line, then the ncu launch list + one full capture.
# usage (from the authoring container):  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh r1a'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $out/gpu_$tag.txt
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -5 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke_$tag.log
tail -2 $out/smoke_$tag.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cat $out/bench_$tag.json; tail -3 $out/bench_$tag.err
timeout 300 python bench.py --workload uniform_64k_direct --steps 10 --warmup 3 --no-cpu-baseline > $out/bench64k_$tag.json 2>> $out/bench_$tag.err
cat $out/bench64k_$tag.json
if [ "$2" != "noncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $out/plain_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
  echo "ncu launches rc=$?"
  CMD2="python bench.py --workload uniform_64k_direct --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD2 > $out/plain64k_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:direct_ -s 3 -c 2 -f -o $out/prof_direct_$tag $CMD2 > $out/ncu_full_$tag.log 2>&1
  echo "ncu full rc=$?"
fi
