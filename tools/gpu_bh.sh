#!/bin/bash
# BH bring-up on the GPU: parity tests (fail-fast off so one call shows every failure) + a timing print.
tag=${1:-bh}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_bh.py -q -m gpu --timeout 600 > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -40 $out/pytest_bh_$tag.log
[ "$2" = "timing" ] && { timeout 600 python tools/bh_timing.py > $out/bh_timing_$tag.log 2>&1; tail -30 $out/bh_timing_$tag.log; }
