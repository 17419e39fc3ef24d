#!/bin/bash
# Round 2, call G (1 GPU): warp-specialised walk - parity tests on the new kernel, then A/B timing against the single-warp walk.
tag=${1:-r2g}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 $out/pytest_bh_$tag.log
for mode in 1 0; do
  echo "NBODY_WALK=$mode" >> $out/bh_walk_ab_$tag.log
  NBODY_WALK=$mode timeout 300 python tools/bh_timing.py 1048576,16777216 walk >> $out/bh_walk_ab_$tag.log 2>&1
done
cat $out/bh_walk_ab_$tag.log
