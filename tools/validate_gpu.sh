#!/usr/bin/env bash
# GPU validation in as few GPU-minutes as possible (run under gpurun; pass the number of GPUs of the box):
#   gpurun --timeout 1500 -- 'bash tools/validate_gpu.sh 1 > gpurun_out/validate1.log 2>&1'            (1 GPU)
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/validate_gpu.sh 2 > gpurun_out/validate2.log 2>&1'   (2 GPUs)
# N=1: kernel numerics + GEMM perf table, the whole gpu test-suite, bench, kernel-variant A/B, ncu step breakdown.
# N=2: multi-GPU tests, bench, fused-TP / fused-MoE micro-benches, env-gated switches, reducer A/B.   N=4: + convergence.
# N=8: bench + BASELINE configs #3-#5.   Box-to-box variance is ~10 %: compare numbers from ONE call only.
set -uo pipefail
N="${1:-1}"
mkdir -p gpurun_out
# GPU-minutes are charged x N: everything that needs ONE GPU runs only in the N=1 call, the N>=2 calls run only what
# needs several GPUs.  Typical session: `validate_gpu.sh 1` (1-GPU box), then `validate_gpu.sh 2`, later `validate_gpu.sh 8`.
if [ "$N" -eq 1 ]; then
  echo "== gemm_check (correctness of every layout / epilogue / CTA-pair variant + perf vs cuBLAS)"
  timeout 400 python tools/gemm_check.py epi > gpurun_out/gemm_check.log 2>&1; grep -c "^OK" gpurun_out/gemm_check.log; grep "FAIL" gpurun_out/gemm_check.log | head; grep "cta1:" gpurun_out/gemm_check.log
  echo "== pytest -m gpu (multi-GPU tests skip themselves)"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
else
  echo "== pytest -m gpu: the multi-GPU tests"
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hybrid.py -m gpu -x -q 2>&1 | tail -4
fi
for n in 1 2 4 8; do
  if [ "$n" -eq "$N" ]; then
    echo "== bench N=$n"
    python bench.py --gpus "$n" --steps 10 --warmup 3 | tee "gpurun_out/bench_${n}gpu.json" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], 'GPUs', round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'tok/s e2e', round(d['e2e']['ms_per_step'], 2), d['clocks'])
"
  fi
done
if [ "$N" -eq 2 ]; then
  echo "== fused TP kernels vs NCCL + GEMM (T=2)"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/tp_bench.py 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['op'], 'fused', round(d['fused_ms'] * 1e3, 1), 'us  nccl+gemm', round(d['nccl_plus_gemm_ms'] * 1e3, 1), ' gemm only', round(d['gemm_only_ms'] * 1e3, 1), ' frac of roofline', round(d['fused_frac_of_roofline'], 2))
"
  echo "== fused MoE layer vs reference-style layer (T=2; config #4 wants T=8: gpurun --gpus 8)"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tools/moe_bench.py 2>&1 | grep "^{" | cut -c1-400
  echo "== candidate switches (written without GPU access; flip the defaults if they pass and win)"
  PIPEGOOSE_B200_FUSED_LM_HEAD=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "tp2_bloom" 2>&1 | tail -2
  for sw in "PIPEGOOSE_B200_FUSED_LM_HEAD=1" "PIPEGOOSE_B200_LNBWD_TO_STAGE=1" "PIPEGOOSE_B200_RS_FUSED_REDUCE=1" "PIPEGOOSE_B200_NCOMM=8"; do
    echo "-- $sw"
    env $sw python bench.py --gpus 2 --steps 10 --warmup 3 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 2), 'ms/step', 'loss', d['final_loss'])
"
  done
  echo "== TP2 step breakdown (torch profiler, diagnosis only)"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dist_step_profile.py --tp 2 2>&1 | grep -v "^\*\|OMP\|^$\|arn" | head -24
fi
if [ "$N" -eq 4 ]; then
  echo "== convergence: TP2 x DP2 + ZeRO-1 (bf16, fused kernels) next to a single-GPU model"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 examples/convergence_hybrid.py --tp 2 --dp 2 --steps 60 2>&1 | grep "^step\|^loss" | tee gpurun_out/convergence_tp2dp2.txt | tail -5
fi
if [ "$N" -ge 8 ]; then
  echo "== BASELINE.json configs #3-#5 (first runs at full size: each under its own timeout)"
  for cfg in "--model bloom-7b1 --tp 8 --seq-len 2048 --batch-per-gpu 1" \
             "--tp 8 --experts 8" \
             "--model bloom-3b --tp 2 --pp 2 --microbatches 8 --batch-per-gpu 2"; do
    echo "-- bench.py --gpus 8 $cfg"
    timeout 420 python bench.py --gpus 8 --steps 5 --warmup 3 $cfg 2>&1 | grep "^{" | tee -a gpurun_out/bench_8gpu_other_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['model'], d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step', round(d['value']), 'tok/s', 'loss', d['final_loss'])
"
  done
  echo "-- fused TP kernels at bloom-7b1 shapes (T=8)"
  TPB_MODEL=7b1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 tools/tp_bench.py 2>&1 | grep "^{" | cut -c1-300
  echo "-- fused MoE layer (T=8)"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 tools/moe_bench.py 2>&1 | grep "^{" | cut -c1-400
fi
echo "== kernel variants built by tools/build_variants.sh (same Python, different _C_<name>.so): numerics, then A/B"
bench_line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', round(d['ms_per_step'], 2), 'ms/step  loss', d['final_loss'])
"; }
if [ -f pipegoose_b200/_C_pdl.so ] && [ "$N" -eq 1 ]; then
  echo "-- pdl: kernel tests"; PIPEGOOSE_B200_EXT=pdl PIPEGOOSE_B200_PDL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
  echo "-- main 1 GPU";           python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
  echo "-- pdl (attribute off)";  PIPEGOOSE_B200_EXT=pdl python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
  echo "-- pdl (PDL=1)";          PIPEGOOSE_B200_EXT=pdl PIPEGOOSE_B200_PDL=1 python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
fi
if [ -f pipegoose_b200/_C_coresident.so ]; then
  echo "-- coresident: kernel tests + GEMM table"; PIPEGOOSE_B200_EXT=coresident timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
  PIPEGOOSE_B200_EXT=coresident timeout 400 python tools/gemm_check.py epi 2>&1 | grep "cta1:\|FAIL" | head -6
  echo "-- coresident 1 GPU (cost of 152 registers)"; PIPEGOOSE_B200_EXT=coresident python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
  if [ "$N" -ge 2 ]; then
    echo "-- dp2 main (64 x 512-thread CTAs)";            python bench.py --gpus 2 --tp 1 --steps 10 --warmup 3 | bench_line
    echo "-- dp2 coresident, old reducer";                PIPEGOOSE_B200_EXT=coresident python bench.py --gpus 2 --tp 1 --steps 10 --warmup 3 | bench_line
    for ctas in -148 -296; do
      echo "-- dp2 coresident, $ctas small CTAs";         PIPEGOOSE_B200_EXT=coresident PIPEGOOSE_B200_DP_OVERLAP_CTAS=$ctas python bench.py --gpus 2 --tp 1 --steps 10 --warmup 3 | bench_line
    done
  fi
fi
[ "$N" -eq 1 ] || exit 0
echo "== 1-GPU step breakdown (ncu launch list)"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/step_profile.py > /dev/null 2>&1
python tools/step_profile.py --aggregate gpurun_out/step_launches.csv gpurun_out/step_breakdown.json | head -22
