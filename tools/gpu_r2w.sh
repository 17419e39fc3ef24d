#!/bin/bash
tag=${1:-r2w}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_gpu_let.py -q -m gpu -x > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest_bh_$tag.log
for cfg in "2 4" "2 2" "0 4"; do set -- $cfg
  echo "NBODY_WALK=$1 NBODY_WALK_SHARE=$2" | tee -a $out/bh_share_$tag.log
  NBODY_WALK=$1 NBODY_WALK_SHARE=$2 timeout 300 python tools/bh_timing.py 1048576,16777216 walk 2>&1 | grep "theta=" | tee -a $out/bh_share_$tag.log
done
