#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_let_phases.sh N workload tag'
N=${1:-2}; wl=${2:-two_galaxies_4m_bh}; tag=${3:-x}
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --workload $wl --bh-exchange 0 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_${wl}_${N}gpu_$tag.json 2> $out/bench_${wl}_${N}gpu_$tag.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('$out/bench_${wl}_${N}gpu_$tag.json').read().strip().splitlines()[-1])
print(d['config']['name'], d['value'], 'steps/s e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['phases_ms_per_step_max_over_ranks'])
print('max', d.get('domain_split_phases_ms_per_step_max_over_ranks'), 'min', d.get('domain_split_phases_ms_per_step_min_over_ranks'), 'let', d['let_points_max'], 'bodies', d['bodies_per_rank'], 'launches', d['gpu_launches'], d['steps'])
PY
tail -3 $out/bench_${wl}_${N}gpu_$tag.err
