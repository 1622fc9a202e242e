#!/usr/bin/env bash
# Where does the data-parallel overhead go?  (2 GPUs, ~4 min)
set -uo pipefail
mkdir -p gpurun_out
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'])
"; }
echo "== fixed tests"
timeout 300 python -m pytest tests/test_gpu_nvlink_engine.py -q -k "workspace" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_zz_gpu_gpt2.py tests/test_zz_gpu_padded_attention.py -q 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== phases + kernel table: 1 GPU"
timeout 200 $TR --nproc-per-node 1 --master-port 29551 tools/dist_step_profile.py --tp 1 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -34
cp gpurun_out/dist_profile_tp1dp1.json gpurun_out/dist_profile_tp1dp1_r2.json
echo "== phases + kernel table: DP2 (co-resident reducer, defaults)"
timeout 200 $TR --nproc-per-node 2 --master-port 29552 tools/dist_step_profile.py --tp 1 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -40
cp gpurun_out/dist_profile_tp1dp2.json gpurun_out/dist_profile_tp1dp2_r2.json
echo "== DP2 bench: defaults / NCCL reducer / no ZeRO all-gather overlap knobs"
timeout 200 python bench.py --gpus 2 --tp 1 --steps 8 --warmup 3 --no-self-check | line
PIPEGOOSE_B200_FUSED_DP=0 timeout 200 python bench.py --gpus 2 --tp 1 --steps 8 --warmup 3 --no-self-check | line
echo "== done"
