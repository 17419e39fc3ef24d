#!/bin/bash
# Round 2 multi-GPU call: NCCL parity check, the default bench line (direct + bh object) and the reference arm under torchrun.
# usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh N tag'
N=${1:-2}; tag=${2:-r2m}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/multi_gpu_check.py > $out/multi_gpu_check_${N}gpu_$tag.log 2>&1; echo "multi check rc=$?"
grep -E "method=|MULTI-GPU|Error|error" $out/multi_gpu_check_${N}gpu_$tag.log | tail -8
timeout 900 $TR --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_${N}gpu_$tag.json 2> $out/bench_${N}gpu_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$out/bench_${N}gpu_$tag.json').read().strip().splitlines()[-1])
    print('direct', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
    b=d.get('bh')
    if b: print('bh', b['config']['name'], b['value'], 'steps/s e2e', b['e2e']['value'], 'ms', b['ms_per_step'], b['phases_ms_per_step_max_over_ranks'], 'let', b['let_points_max'], 'bodies', b['bodies_per_rank'], 'clk', b['clocks'])
except Exception as e:
    print('parse failed', e)
PY
tail -5 $out/bench_${N}gpu_$tag.err
timeout 600 $TR --master-port 29613 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-extras > $out/bench_ref_${N}gpu_$tag.json 2>> $out/bench_${N}gpu_$tag.err; echo "ref rc=$?"
cut -c1-400 $out/bench_ref_${N}gpu_$tag.json
