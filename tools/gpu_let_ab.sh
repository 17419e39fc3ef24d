#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_let_ab.sh <tag> N <workload> <steps> <variants...>' ; variants: let serial repl direct
tag=$1; N=$2; wl=$3; K=${4:-10}; shift 4; out=gpurun_out; mkdir -p $out
one() { f=$out/ab_$1_g${N}_$tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus $N --steps $K --warmup 3 --e2e-steps 1 ${@:2} > $f.json 2> $f.err
  python -c "
import json; d=json.loads(open('$f.json').read().strip().splitlines()[-1]); print('$1', 'gpus', d['n_gpus'], d['value'], d['unit'], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms_per_step'].items()}, 'e2e', d['e2e']['value'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 $f.err; }
for v in "$@"; do
  case $v in
    let) one let --workload $wl --bh-exchange 0;;
    serial) NBODY_LET_NO_OVERLAP=1 one let_serial --workload $wl --bh-exchange 0;;
    repl) one replicated --workload $wl --bh-exchange 1;;
    direct) one direct1m;;
  esac
done
