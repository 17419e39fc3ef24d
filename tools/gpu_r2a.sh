#!/bin/bash
# Round 2, call A (1 GPU): full GPU test suite (incl. loop-back LET, config 3 / 5, UE adapter), smoke, default bench line
# (direct + bh object), parameter landscape of the walk, compute-sanitizer memcheck.
tag=${1:-r2a}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $out/gpu_$tag.txt
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -30 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke_$tag.log
tail -2 $out/smoke_$tag.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cat $out/bench_$tag.json | cut -c1-6000; tail -3 $out/bench_$tag.err
NBODY_NO_EQUAL_MASS=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bh > $out/bench_generalmass_$tag.json 2>> $out/bench_$tag.err
python -c "import json; d=json.load(open('$out/bench_generalmass_$tag.json')); print('general-mass kernel:', d['value'], d['roofline']['frac'], d['roofline']['jsplit'])"
timeout 600 python tools/bh_timing.py 1048576 sweep > $out/bh_sweep_$tag.log 2>&1; tail -20 $out/bh_sweep_$tag.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py > $out/sanitizer_memcheck_$tag.log 2>&1; echo "memcheck rc=$?"
tail -5 $out/sanitizer_memcheck_$tag.log
