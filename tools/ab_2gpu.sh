#!/usr/bin/env bash
# Short 2-GPU follow-up (one gpurun call): NVLink-engine tests, co-resident reducer A/B on the DP2 step, TP2 re-check.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/ab_2gpu.sh > gpurun_out/ab_2gpu.log 2>&1'
set -uo pipefail
mkdir -p gpurun_out
S="--steps 8 --warmup 3"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); n = d.get('numerics') or {}
        print('   ', d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'],
              ' numerics_ok', d.get('numerics_ok'), n.get('max_rel_err_loss'), n.get('error', ''), d['clocks']['reasons'])
"; }
echo "== NVLink engine tests"
timeout 600 python -m pytest tests/test_gpu_nvlink_engine.py -q -s 2>&1 | grep -v "^W0\|warn" | grep "rank\|passed\|failed\|Error\|assert" | tail -20
echo "== DP2 (no TP): bucketed reducer with 64 big CTAs (default) vs co-resident small CTAs"
timeout 300 python bench.py --gpus 2 --tp 1 $S | line
for ctas in -148 -296 -592; do
  echo "-- PIPEGOOSE_B200_DP_OVERLAP_CTAS=$ctas"
  PIPEGOOSE_B200_DP_OVERLAP_CTAS=$ctas timeout 300 python bench.py --gpus 2 --tp 1 $S --no-self-check | line
done
echo "-- PIPEGOOSE_B200_DP_OVERLAP_CTAS=-296 PIPEGOOSE_B200_DP_TAIL_CTAS=-592"
PIPEGOOSE_B200_DP_OVERLAP_CTAS=-296 PIPEGOOSE_B200_DP_TAIL_CTAS=-592 timeout 300 python bench.py --gpus 2 --tp 1 $S --no-self-check | line
echo "== TP2 (defaults now: fused lm_head, LN backward to stage)"
timeout 300 python bench.py --gpus 2 $S | tee gpurun_out/bench_2gpu_tp2_v2.json | line
PIPEGOOSE_B200_CE_IN_EPILOGUE=1 timeout 300 python bench.py --gpus 2 $S --no-self-check | line
echo "-- peer-memory small collectives (no NCCL kernel in the TP step): numerics, then the step"
PIPEGOOSE_B200_TP_PEER_COLLECTIVES=1 timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "tp2_bloom" 2>&1 | tail -2
PIPEGOOSE_B200_TP_PEER_COLLECTIVES=1 timeout 300 python bench.py --gpus 2 $S | line
PIPEGOOSE_B200_TP_PEER_COLLECTIVES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dist_step_profile.py --tp 2 2>&1 | grep -i "nccl\|fwd" | head -8
echo "== 1 GPU"
timeout 300 python bench.py --gpus 1 $S | line
echo "== MoE EP2 end to end (rehearsal of config #4) + PP2 (rehearsal of config #5's engine)"
timeout 300 python bench.py --gpus 2 --tp 2 --experts 4 --steps 4 --warmup 3 2>&1 | grep "^{\|Error" | line
timeout 300 python bench.py --gpus 2 --tp 1 --pp 2 --microbatches 4 --steps 4 --warmup 3 2>&1 | grep "^{\|Error" | line
echo "== done"
