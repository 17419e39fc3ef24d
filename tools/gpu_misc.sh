#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_bh.py -q -m gpu > $out/pytest_bh_$tag.log 2>&1; tail -2 $out/pytest_bh_$tag.log
timeout 300 python tools/bh_timing.py 1048576,16777216 2>&1 | grep -v "theta=0.35" | tail -4
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_d_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:direct_packed -s 3 -c 1 -f -o $out/prof_direct1m_$tag $CMD > $out/ncu_full_d_$tag.log 2>&1
echo "ncu rc=$?"
