#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_let_trace.sh <tag> N <workload> [events|trace]'
tag=$1; N=$2; wl=$3; mode=${4:-trace}; out=gpurun_out; mkdir -p $out
if [ $mode = events ]; then export NBODY_LET_EVENTS=1; else export NBODY_LET_TRACE=1; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus $N --workload $wl --steps 3 --warmup 3 --e2e-steps 1 --bh-exchange 0 > $out/let_trace_$tag.json 2> $out/let_trace_$tag.err
echo rc=$?
grep "^\[let" $out/let_trace_$tag.err | grep "step [5]\]" | sort -s -k3,3n | head -80
python -c "
import json; d=json.loads(open('$out/let_trace_$tag.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'])"
