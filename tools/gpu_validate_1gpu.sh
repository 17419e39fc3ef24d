#!/bin/bash
# Round 2, call F (1 GPU): full GPU test suite, smoke, default bench line (direct + bh + 16M baseline).
tag=${1:-r2f}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -15 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke_$tag.log
timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
cat $out/bench_$tag.json | cut -c1-9000; tail -3 $out/bench_$tag.err
timeout 600 python bench.py --impl reference > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err; echo "ref rc=$?"
cat $out/bench_ref_$tag.json | cut -c1-3000
