#!/bin/bash
# Per-phase host-clock trace of the domain-split step (NBODY_LET_TRACE=1 synchronises after every phase).  usage: gpurun --gpus N -- 'bash tools/gpu_let_trace.sh N tag'
N=${1:-4}; tag=${2:-trace}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NBODY_LET_TRACE=${3:-1} timeout 600 $TR --master-port 29621 bench.py --gpus $N --workload two_galaxies_16m_bh --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/let_trace_${N}gpu_$tag.json 2> $out/let_trace_${N}gpu_$tag.err
echo "rc=$?"; grep "let rank 0 step\|let rank 3 step" $out/let_trace_${N}gpu_$tag.err | head -40
