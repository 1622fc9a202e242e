#!/usr/bin/env bash
# Round-2 single-GPU validation (one gpurun call):
#   gpurun --timeout 1500 -- 'bash tools/validate_1gpu.sh > gpurun_out/validate_1gpu.log 2>&1'
# kernel tests (incl. the cross-entropy statistics in the lm_head epilogue), bench A/B of the single-GPU switches, the
# transformers-model path, compute-sanitizer (memcheck / racecheck / synccheck) on the kernel tests, ncu captures.
set -uo pipefail
mkdir -p gpurun_out
S="--steps 10 --warmup 3"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'], ' mfu', round(d.get('mfu_vs_measured_sustained') or 0, 3), d['clocks'])
"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== pytest -m gpu (1-GPU tests)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench: default, CE statistics in the lm_head epilogue, transformers model"
python bench.py --gpus 1 $S | tee gpurun_out/bench_1gpu.json | line
PIPEGOOSE_B200_CE_IN_EPILOGUE=1 python bench.py --gpus 1 $S | line
python bench.py --gpus 1 $S --hf | line
PIPEGOOSE_B200_CE_IN_EPILOGUE=1 python bench.py --gpus 1 $S --hf | line
python bench.py --gpus 1 $S | line
echo "== compute-sanitizer on the kernel tests"
for tool in memcheck racecheck synccheck; do
  echo "-- $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 --error-exitcode 1 \
      python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or layernorm or cross_entropy or lm_head or embedding or adam" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
echo "== ncu: lm_head GEMM with the CE epilogue; step launch list"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 1 -c 2 -o gpurun_out/ncu_lm_head_ce -f python tools/profile_kernels.py lm_head_ce > /dev/null 2>&1; echo "rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_r2.csv python tools/step_profile.py > /dev/null 2>&1
python tools/step_profile.py --aggregate gpurun_out/step_launches_r2.csv gpurun_out/step_breakdown_r2.json | head -24
echo "== done"
