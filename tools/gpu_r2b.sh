#!/bin/bash
# Round 2, call B (1 GPU): new walk kernel - parity tests, timing landscape, ncu of the walk.
tag=${1:-r2b}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py tests/test_gpu_multi.py tests/test_ue4_adapter.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 $out/pytest_bh_$tag.log
timeout 600 python tools/bh_timing.py 1048576 sweep > $out/bh_sweep_$tag.log 2>&1; cat $out/bh_sweep_$tag.log
timeout 300 python tools/bh_timing.py 4194304,16777216 > $out/bh_timing_$tag.log 2>&1; cat $out/bh_timing_$tag.log
CMD="python bench.py --workload plummer_1m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_bh_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bh_walk_group" -s 4 -c 2 -f -o $out/prof_walk_$tag $CMD > $out/ncu_full_walk_$tag.log 2>&1
echo "ncu full rc=$?"; tail -2 $out/ncu_full_walk_$tag.log | cut -c1-200
