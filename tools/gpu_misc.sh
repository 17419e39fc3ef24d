#!/bin/bash
tag=${1:-x}; out=gpurun_out; mkdir -p $out
CMD="python bench.py --workload two_galaxies_16m_bh --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > $out/plain_16m_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $out/launches_16m_$tag.csv $CMD > $out/ncu_launches_16m_$tag.log 2>&1
echo "ncu rc=$?"; tail -1 $out/plain_16m_$tag.log | cut -c1-400
