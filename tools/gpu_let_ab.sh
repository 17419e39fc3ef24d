#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_let_ab.sh <tag> N <workload> <steps>' : LET overlap on/off, replicated
tag=$1; N=$2; wl=$3; K=${4:-10}; out=gpurun_out; mkdir -p $out
one() { f=$out/ab_$1_$tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus $N --workload $wl --steps $K --warmup 3 --e2e-steps 1 ${@:2} > $f.json 2> $f.err
  python -c "
import json; d=json.loads(open('$f.json').read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms_per_step'].items()})" || tail -5 $f.err; }
one let_overlap --bh-exchange 0
NBODY_LET_NO_OVERLAP=1 one let_serial --bh-exchange 0
one let_overlap2 --bh-exchange 0
one replicated --bh-exchange 1
