#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi2.sh <tag> N <workload> [check]'  : BH both exchange modes (+ direct) at N GPUs
tag=$1; N=$2; wl=$3
out=gpurun_out; mkdir -p $out
run() { # name, extra args
  f=$out/bench_$1_g${N}_$tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N ${@:2} > $f.json 2> $f.err
  echo "bench $1 gpus=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","unit","n_gpus","ms_per_step","phases_ms_per_step")}, "e2e", d["e2e"]["value"], "inter/step", d["roofline"].get("interactions_per_step"))
except Exception as e:
    print("no json", e); print(open("$f.err").read()[-2500:])
PY
}
if [ "$4" = "check" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 tools/multi_gpu_check.py 2>&1 | grep -v "^W10\|^\[W\|NCCL version" | tail -8
fi
run ${wl}_let --workload $wl --steps 10 --warmup 3 --e2e-steps 1 --bh-exchange 0
run ${wl}_repl --workload $wl --steps 10 --warmup 3 --e2e-steps 1 --bh-exchange 1
[ "$5" = "direct" ] && run direct1m --steps 5 --warmup 3 --e2e-steps 1
