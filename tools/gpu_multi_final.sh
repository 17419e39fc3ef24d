#!/bin/bash
# Round 2 final multi-GPU call: NCCL parity check, the default bench line (direct + bh object), the Barnes-Hut workload alone with
# more NCCL point-to-point channels (A/B), the reference arm.   usage: gpurun --gpus N -- 'bash tools/gpu_multi_final.sh N tag'
N=${1:-8}; tag=${2:-r2fin}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/multi_gpu_check.py > $out/multi_gpu_check_${N}gpu_$tag.log 2>&1; echo "multi check rc=$?"
grep -E "method=|MULTI-GPU|Error|error" $out/multi_gpu_check_${N}gpu_$tag.log | tail -8
timeout 900 $TR --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_${N}gpu_$tag.json 2> $out/bench_${N}gpu_$tag.err; echo "bench rc=$?"
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if d.get('metric', '').startswith('all-pairs'):
        print('direct', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
        d = d.get('bh')
    if d:
        print('bh', d['config']['name'], round(d['value'], 1), 'steps/s, e2e', round(d['e2e']['value'], 1), d['e2e'].get('ms_per_call_max_over_ranks'), 'ms', round(d['ms_per_step'], 3),
              d.get('domain_split_phases_ms_per_step_max_over_ranks'), 'let', d['let_points_max'], 'clk', d['clocks'])
except Exception as e:
    print('parse failed', e)
PY
}
summ $out/bench_${N}gpu_$tag.json; tail -3 $out/bench_${N}gpu_$tag.err
NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 timeout 600 $TR --master-port 29614 bench.py --gpus $N --workload two_galaxies_16m_bh --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_bh_p2pch_${N}gpu_$tag.json 2> $out/bench_bh_p2pch_${N}gpu_$tag.err; echo "bh p2p-channels rc=$?"
summ $out/bench_bh_p2pch_${N}gpu_$tag.json
timeout 600 $TR --master-port 29613 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-extras > $out/bench_ref_${N}gpu_$tag.json 2>> $out/bench_${N}gpu_$tag.err; echo "ref rc=$?"
cut -c1-300 $out/bench_ref_${N}gpu_$tag.json
