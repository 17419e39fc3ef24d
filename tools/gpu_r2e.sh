#!/bin/bash
tag=${1:-r2e}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py tests/test_gpu_multi.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 $out/pytest_bh_$tag.log
timeout 600 python tools/bh_timing.py 1048576,4194304,16777216 > $out/bh_timing_$tag.log 2>&1; cat $out/bh_timing_$tag.log
CMD="python bench.py --workload two_galaxies_16m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_16m_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 120 --csv --log-file $out/launches_16m_$tag.csv $CMD > $out/ncu_launches_16m_$tag.log 2>&1
echo "ncu launches rc=$?"
