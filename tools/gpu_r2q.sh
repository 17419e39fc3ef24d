#!/bin/bash
tag=${1:-r2q}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_let.py -q -m gpu -x > $out/pytest_let_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_let_$tag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_let_loopback_$tag.csv python tools/let_probe.py 16777216 8 4 > $out/ncu_let_$tag.log 2>&1
echo "ncu rc=$?"; tail -5 $out/ncu_let_$tag.log
