#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> N [workloads...]'
tag=${1:-m}; N=${2:-2}; shift; shift
WL=${@:-plummer_1m_direct plummer_1m_bh}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -2
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > $out/pytest_multi_$tag.log 2>&1; echo "pytest rc=$?"
tail -12 $out/pytest_multi_$tag.log
for wl in $WL; do
for n in 1 $N; do
  f=$out/bench_${wl}_g${n}_$tag
  if [ $n = 1 ]; then timeout 900 python bench.py --gpus 1 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $f.json 2> $f.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --workload $wl --steps 5 --warmup 3 --e2e-steps 1 > $f.json 2> $f.err; fi
  echo "bench $wl gpus=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$f.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","unit","n_gpus","ms_per_step","phases_ms_per_step")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("no json", e); print(open("$f.err").read()[-1500:])
PY
done
done
