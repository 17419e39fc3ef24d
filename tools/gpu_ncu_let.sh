#!/bin/bash
# ncu --set full of one K9 kernel in the 8-rank loop-back run on ONE GPU.  usage: gpu_ncu_let.sh TAG KERNEL_REGEX [SKIP]
tag=${1:-r2p}; kern=${2:-let_export}; skip=${3:-12}
out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$kern -s $skip -c 1 -f -o $out/prof_let_$tag python tools/let_probe.py 16777216 8 3 > $out/ncu_let_full_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 $out/ncu_let_full_$tag.log
