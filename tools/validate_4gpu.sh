#!/usr/bin/env bash
# 4-GPU box: multi-GPU suites that need 4 GPUs, the TP2 x DP2 bench, small dress rehearsals of the 8-GPU configs, and the
# >= 300-step convergence curves of both arms (one gpurun call):
#   gpurun --gpus 4 --timeout 1500 -- 'bash tools/validate_4gpu.sh > gpurun_out/validate_4gpu.log 2>&1'
set -uo pipefail
mkdir -p gpurun_out
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if 'unavailable' in d: print('   ', d); continue
        n = d.get('numerics') or {}
        print('   ', d.get('impl'), d['config']['model'], d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'tok/s',
              ' loss', d['final_loss'], ' numerics_ok', d.get('numerics_ok'), n.get('max_rel_err_loss'), n.get('error', ''), d.get('note', ''))
"; }
echo "== multi-GPU suites (4-GPU cases included)"
timeout 900 python -m pytest tests/test_gpu_nvlink_engine.py tests/test_gpu_multi.py tests/test_gpu_hybrid.py -q 2>&1 | tail -6
echo "== bench TP2 x DP2 + ZeRO-1 (with self-check); then the round-1 reducer configuration (64 / 96 big CTAs) for comparison"
timeout 300 python bench.py --gpus 4 --steps 10 --warmup 3 | tee gpurun_out/bench_4gpu.json | line
PIPEGOOSE_B200_DP_OVERLAP_CTAS=64 PIPEGOOSE_B200_DP_TAIL_CTAS=96 timeout 300 python bench.py --gpus 4 --steps 10 --warmup 3 --no-self-check | line
echo "== TP2 x DP2 step: phases and kernel table (torch profiler, diagnosis only)"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29539 tools/dist_step_profile.py --tp 2 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -30
echo "== dress rehearsal of configs #3-#5 at 4 GPUs"
timeout 300 python bench.py --gpus 4 --steps 3 --warmup 3 --model bloom-7b1 --tp 4 --seq-len 2048 --batch-per-gpu 1 2>&1 | grep "^{\|Error" | line
timeout 300 python bench.py --gpus 4 --steps 3 --warmup 3 --model bloom-3b --tp 2 --pp 2 --microbatches 8 --batch-per-gpu 2 2>&1 | grep "^{\|Error" | line
echo "== convergence, 300 steps: TP2 x DP2 + ZeRO-1 (ours, reference)"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 4 --master-port 29541 tools/convergence_gpu.py --gpus 4 --steps 300 --out gpurun_out/convergence_tp2dp2_zero1_b200.txt 2>&1 | grep "^step\|^{" | tail -4
timeout 600 $TR --nproc-per-node 4 --master-port 29542 tools/convergence_gpu.py --gpus 4 --steps 300 --impl reference --out gpurun_out/convergence_tp2dp2_zero1_reference_b200.txt 2>&1 | grep "^step\|^{" | tail -4
echo "== done"
