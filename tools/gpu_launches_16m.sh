#!/bin/bash
# ncu launch list (duration + DRAM bytes) of the 16M-body Barnes-Hut step on one GPU (after the plain command exited 0).
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
CMD="python bench.py --workload two_galaxies_16m_bh --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $CMD > $out/plain_16m_$tag.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 80 --csv --log-file $out/launches_16m_$tag.csv $CMD > $out/ncu_launches_16m_$tag.log 2>&1
echo "rc=$?"
