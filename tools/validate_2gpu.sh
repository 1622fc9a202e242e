#!/usr/bin/env bash
# Round-2 validation on a 2-GPU box (one gpurun call):
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/validate_2gpu.sh > gpurun_out/validate_2gpu.log 2>&1'
# NVLink engine pieces (VMM + multicast workspace, in-kernel gradient reduce-scatter), the existing multi-GPU suites,
# then A/B of the switches on the TP2 and DP2 step.  Everything runs under its own timeout; in-kernel spins trap after ~30 s.
set -uo pipefail
mkdir -p gpurun_out
S="--steps 8 --warmup 3"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); n = d.get('numerics') or {}
        print('   ', d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'],
              ' numerics_ok', d.get('numerics_ok'), n.get('max_rel_err_loss'), n.get('rel_err_param_checksum'), n.get('error', ''), d['clocks']['reasons'])
"; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
nvidia-smi topo -m | head -6
echo "== NVLink engine: workspace / multicast / gradient kernels / ZeRO-1 in-kernel reduce-scatter"
timeout 600 python -m pytest tests/test_gpu_nvlink_engine.py -x -q -s 2>&1 | grep -v "^W0\|warn" | tail -25
echo "== existing multi-GPU suites"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hybrid.py -q 2>&1 | tail -8
echo "== fused lm_head numerics"
PIPEGOOSE_B200_FUSED_LM_HEAD=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "tp2_bloom" 2>&1 | tail -2
echo "== does NCCL itself use NVLS here?"
NCCL_DEBUG=INFO timeout 200 python bench.py --gpus 2 --steps 2 --warmup 3 --no-self-check 2>&1 | grep -i "nvls" | head -5
echo "== TP2 step: baseline (with numerics self-check), then switches"
timeout 300 python bench.py --gpus 2 $S | tee gpurun_out/bench_2gpu_tp2.json | line
for sw in "PIPEGOOSE_B200_FUSED_LM_HEAD=1" "PIPEGOOSE_B200_LNBWD_TO_STAGE=1" "PIPEGOOSE_B200_RS_FUSED_REDUCE=1" "PIPEGOOSE_B200_NCOMM=8" \
          "PIPEGOOSE_B200_FUSED_LM_HEAD=1 PIPEGOOSE_B200_LNBWD_TO_STAGE=1" "PIPEGOOSE_B200_PDL=0" "PIPEGOOSE_B200_SYMM=ipc"; do
  echo "-- $sw"
  env $sw timeout 300 python bench.py --gpus 2 $S --no-self-check | line
done
echo "== TP2 with the transformers model (--hf)"
timeout 300 python bench.py --gpus 2 $S --hf | line
echo "== DP2 step (no TP): in-kernel gradient reduce-scatter on / off, NVLS on / off"
timeout 300 python bench.py --gpus 2 --tp 1 $S | tee gpurun_out/bench_2gpu_dp2.json | line
for sw in "PIPEGOOSE_B200_DP_INLINE_RS=0" "PIPEGOOSE_B200_DP_INLINE_RS=0 PIPEGOOSE_B200_NVLS=0" "PIPEGOOSE_B200_NVLS=0" "PIPEGOOSE_B200_DP_INLINE_SCALAR=1"; do
  echo "-- $sw"
  env $sw timeout 300 python bench.py --gpus 2 --tp 1 $S --no-self-check | line
done
echo "== 1 GPU on this box (reference point for the scaling ratios)"
timeout 300 python bench.py --gpus 1 $S | line
echo "== fused TP kernels vs NCCL + GEMM (T=2)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/tp_bench.py 2>&1 | grep "^{" | tee gpurun_out/tp_bench_T2.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['op'], 'fused', round(d['fused_ms'] * 1e3, 1), 'us  nccl+gemm', round(d['nccl_plus_gemm_ms'] * 1e3, 1), ' gemm only', round(d['gemm_only_ms'] * 1e3, 1), ' frac of roofline', round(d['fused_frac_of_roofline'], 2))
"
echo "== fused MoE layer vs reference-style layer (T=2)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tools/moe_bench.py 2>&1 | grep "^{" | tee gpurun_out/moe_bench_T2.jsonl | cut -c1-600
echo "== TP2 step breakdown (torch profiler, diagnosis only)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dist_step_profile.py --tp 2 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -30
cp gpurun_out/dist_profile_tp2dp1.json gpurun_out/dist_profile_tp2dp1_r2.json 2>/dev/null
echo "== done"
