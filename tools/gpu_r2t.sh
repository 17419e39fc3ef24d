#!/bin/bash
tag=${1:-r2t}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_bh_$tag.log
timeout 600 python tools/bh_timing.py 1048576,2097152,16777216 > $out/bh_timing_$tag.log 2>&1; grep "theta=0.25" $out/bh_timing_$tag.log
