#!/bin/bash
tag=${1:-r2s}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_bh_$tag.log
timeout 600 python tools/bh_timing.py 1048576,16777216 > $out/bh_timing_$tag.log 2>&1; grep "theta=0.25" $out/bh_timing_$tag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_let_loopback_$tag.csv python tools/let_probe.py 16777216 8 4 > $out/ncu_let_$tag.log 2>&1
echo "ncu rc=$?"; python tools/launch_agg.py $out/launches_let_loopback_$tag.csv 16 | head -8
