#!/bin/bash
# BH on the GPU: parity tests, optional timing sweep / bench line.
tag=${1:-bh}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_bh.py tests/test_gpu_multi.py -q -m gpu --timeout 600 > $out/pytest_bh_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 $out/pytest_bh_$tag.log
[ "$2" = "timing" ] && { timeout 600 python tools/bh_timing.py > $out/bh_timing_$tag.log 2>&1; tail -30 $out/bh_timing_$tag.log; }
if [ "$2" = "bench" ]; then
  for wl in plummer_1m_bh two_galaxies_2m_bh; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out/bench_${wl}_$tag.json 2> $out/bench_${wl}_$tag.err; echo "bench rc=$?"
  python -c "
import json; d=json.loads(open('$out/bench_${wl}_$tag.json').read().strip().splitlines()[-1]); print('$wl', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])" || tail -5 $out/bench_${wl}_$tag.err
  done
fi
