#!/usr/bin/env bash
# BASELINE.json configs #3, #4, #5 on one 8-GPU box, this library next to the reference arm (one gpurun call, ~8 min):
#   gpurun --gpus 8 --timeout 560 -- 'bash tools/validate_8gpu.sh > gpurun_out/validate_8gpu.log 2>&1'
# GPU-minutes are charged x8: 3 timed steps per run, every command under its own timeout, no self-check here (it ran at
# 2 and 4 GPUs and runs again in the driver's scaling bench).
set -uo pipefail
mkdir -p gpurun_out
S="--steps 3 --warmup 3 --no-self-check"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'unavailable' in d: print('   ', d); continue
        print('   ', d.get('impl'), d['config']['model'], d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'tok/s  e2e', round(d['e2e']['value']),
              ' loss', d['final_loss'], d.get('note', ''), (d.get('clocks') or {}).get('reasons'))
"; }
run() {  # run <tag> <timeout ours> <timeout ref> <bench args...>: both arms, JSON lines appended to gpurun_out/bench_8gpu_configs.jsonl
  tag=$1; to=$2; tr=$3; shift 3
  echo "-- $tag: ours"
  timeout $to python bench.py --gpus 8 $S "$@" 2> gpurun_out/${tag}_ours.err | grep "^{" | tee -a gpurun_out/bench_8gpu_configs.jsonl | line
  grep -i "error\|Traceback" gpurun_out/${tag}_ours.err | tail -2
  echo "-- $tag: reference"
  timeout $tr python bench.py --impl reference --gpus 8 $S "$@" 2> gpurun_out/${tag}_ref.err | grep "^{" | tee -a gpurun_out/bench_8gpu_configs.jsonl | line
  grep -i "error\|Traceback" gpurun_out/${tag}_ref.err | tail -2
}
echo "== config #2 (headline): bloom-560m TP2 x DP4 + ZeRO-1, NVLS (multimem) reducer / all-gather on (default at dp = 4) and off"
timeout 120 python bench.py --gpus 8 --steps 6 --warmup 3 --no-self-check | tee gpurun_out/bench_8gpu_headline.json | line
PIPEGOOSE_B200_NVLS_REDUCE=0 PIPEGOOSE_B200_NVLS_ALLGATHER=0 timeout 120 python bench.py --gpus 8 --steps 6 --warmup 3 --no-self-check | line
echo "== config #3: bloom-7b1 TP=8 seq 2048"
run cfg3 150 200 --model bloom-7b1 --tp 8 --seq-len 2048 --batch-per-gpu 1
echo "== config #4: bloom-560m + Switch-MoE 8 experts, EP=8, Top-1"
run cfg4 120 160 --tp 8 --experts 8
echo "== config #5: bloom-3b TP2 x PP2 x DP2 + ZeRO-1, 1F1B, 8 micro-batches"
run cfg5 150 180 --model bloom-3b --tp 2 --pp 2 --microbatches 8 --batch-per-gpu 2
echo "== layout check (partitioning/planner.py ranks pure DP8 + ZeRO-1 ahead of the prescribed TP2 x DP4 for bloom-560m: dp 2 measured +7.6 %, tp 2 +12.7 %)"
timeout 120 python bench.py --gpus 8 --tp 1 --steps 6 --warmup 3 --no-self-check | line
echo "== done"
