#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> N'
tag=${1:-m}; N=${2:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_bh.py -q -m gpu --timeout 600 > $out/pytest_multi_$tag.log 2>&1; echo "pytest rc=$?"
tail -25 $out/pytest_multi_$tag.log
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_g1_$tag.json 2> $out/bench_g1_$tag.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 3 --warmup 3 > $out/bench_g${n}_$tag.json 2> $out/bench_g${n}_$tag.err; fi
  echo "bench gpus=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/bench_g${n}_$tag.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","phases_ms_per_step","frac_fp32_peak_per_gpu")}, d["e2e"]["value"])
except Exception as e:
    print("no json", e); print(open("$out/bench_g${n}_$tag.err").read()[-1500:])
PY
done
