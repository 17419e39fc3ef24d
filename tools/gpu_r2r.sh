#!/bin/bash
tag=${1:-r2r}
out=gpurun_out; mkdir -p $out
for v in 0 1; do
NBODY_LET_ACCEPT=$v timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:let_export -c 24 --csv --log-file $out/launches_export_v${v}_$tag.csv python tools/let_probe.py 16777216 8 3 > $out/ncu_let_v${v}_$tag.log 2>&1
echo "variant $v ncu rc=$?"; grep -c let_export $out/launches_export_v${v}_$tag.csv
python tools/launch_agg.py $out/launches_export_v${v}_$tag.csv 1
done
