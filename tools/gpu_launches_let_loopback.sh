#!/bin/bash
# Launch list of the domain-split step at the 8-rank shape (16M bodies, 8 loop-back ranks on ONE GPU): per-kernel durations of K9.
tag=${1:-r2n}
out=gpurun_out; mkdir -p $out
timeout 600 python tools/let_probe.py 16777216 8 4 > $out/let_probe_$tag.log 2>&1 || { tail -5 $out/let_probe_$tag.log; exit 1; }
cat $out/let_probe_$tag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_let_loopback_$tag.csv python tools/let_probe.py 16777216 8 4 > $out/ncu_let_$tag.log 2>&1
echo "ncu rc=$?"
