#!/bin/bash
tag=${1:-r2v}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_bh.py -q -m gpu -x -k "sort or tree or config4" > $out/pytest_sort_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_sort_$tag.log
for n in 1048576 2097152 16777216; do timeout 120 python tools/sort_probe.py $n; done
