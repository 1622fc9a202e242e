#!/usr/bin/env bash
# BASELINE.json configs #3, #4, #5 on one 8-GPU box, this library next to the reference arm (one gpurun call):
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/validate_8gpu.sh > gpurun_out/validate_8gpu.log 2>&1'
# GPU-minutes are charged x8: short runs (4 timed steps), every command under its own timeout.
set -uo pipefail
mkdir -p gpurun_out
S="--steps 4 --warmup 3"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'unavailable' in d: print('   ', d); continue
        n = d.get('numerics') or {}
        print('   ', d.get('impl'), d['config']['model'], d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'tok/s  e2e', round(d['e2e']['value']),
              ' loss', d['final_loss'], ' numerics_ok', d.get('numerics_ok'), n.get('max_rel_err_loss'), n.get('error', ''), d.get('note', ''), (d.get('clocks') or {}).get('reasons'))
"; }
run() {  # run <tag> <bench args...>: both arms, JSON lines appended to gpurun_out/bench_8gpu_configs.jsonl
  tag=$1; shift
  echo "-- $tag: ours"
  timeout 420 python bench.py --gpus 8 $S "$@" 2> gpurun_out/${tag}_ours.err | grep "^{" | tee -a gpurun_out/bench_8gpu_configs.jsonl | line
  tail -3 gpurun_out/${tag}_ours.err | grep -i "error\|Traceback" | head -3
  echo "-- $tag: reference"
  timeout 420 python bench.py --impl reference --gpus 8 $S "$@" 2> gpurun_out/${tag}_ref.err | grep "^{" | tee -a gpurun_out/bench_8gpu_configs.jsonl | line
  tail -3 gpurun_out/${tag}_ref.err | grep -i "error\|Traceback" | head -3
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | head -3
echo "== config #2 (headline): bloom-560m TP2 x DP4 + ZeRO-1 (ours only; the driver runs both arms at round end)"
timeout 300 python bench.py --gpus 8 --steps 8 --warmup 3 | tee gpurun_out/bench_8gpu_headline.json | line
echo "-- NVLS off (multimem.ld_reduce / multimem.st not used by the dp = 4 reducer and all-gather)"
PIPEGOOSE_B200_NVLS_REDUCE=0 PIPEGOOSE_B200_NVLS_ALLGATHER=0 timeout 300 python bench.py --gpus 8 --steps 8 --warmup 3 --no-self-check | line
echo "== TP2 x DP4 step: phases and kernel table (torch profiler, diagnosis only)"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 tools/dist_step_profile.py --tp 2 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -30
echo "== config #3: bloom-7b1 TP=8 seq 2048"
run cfg3 --model bloom-7b1 --tp 8 --seq-len 2048 --batch-per-gpu 1
echo "== config #4: bloom-560m + Switch-MoE 8 experts, EP=8, Top-1"
run cfg4 --tp 8 --experts 8
echo "== config #5: bloom-3b TP2 x PP2 x DP2 + ZeRO-1, 1F1B, 8 micro-batches"
run cfg5 --model bloom-3b --tp 2 --pp 2 --microbatches 8 --batch-per-gpu 2
echo "== fused TP kernels at bloom-7b1 shapes (T=8) vs NCCL + GEMM"
TPB_MODEL=7b1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 tools/tp_bench.py 2>&1 | grep "^{" | tee gpurun_out/tp_bench_7b1_T8.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['op'], 'fused', round(d['fused_ms'] * 1e3, 1), 'us  nccl+gemm', round(d['nccl_plus_gemm_ms'] * 1e3, 1), ' gemm only', round(d['gemm_only_ms'] * 1e3, 1), ' frac of roofline', round(d['fused_frac_of_roofline'], 2), d.get('bound'))
"
echo "== fused MoE layer (T=8) vs reference-style layer"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 tools/moe_bench.py 2>&1 | grep "^{" | tee gpurun_out/moe_bench_T8.jsonl | cut -c1-700
echo "== done"
