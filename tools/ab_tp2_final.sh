#!/usr/bin/env bash
# Which of the round-2 changes cost the TP2 step its forward millisecond?  (2 GPUs, ~3 min)
set -uo pipefail
mkdir -p gpurun_out
S="--gpus 2 --steps 8 --warmup 3 --no-self-check"
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', d['config']['parallelism'], round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'])
"; }
echo "-- defaults";                                   timeout 150 python bench.py $S | line
echo "-- GEMM with the round-1 register budget";      PIPEGOOSE_B200_EXT=oldreg timeout 150 python bench.py $S | line
echo "-- CE statistics as a separate pass";           PIPEGOOSE_B200_CE_IN_EPILOGUE=0 timeout 150 python bench.py $S | line
echo "-- small collectives on NCCL";                  PIPEGOOSE_B200_TP_PEER_COLLECTIVES=0 timeout 150 python bench.py $S | line
echo "-- all three";                                  PIPEGOOSE_B200_EXT=oldreg PIPEGOOSE_B200_CE_IN_EPILOGUE=0 PIPEGOOSE_B200_TP_PEER_COLLECTIVES=0 timeout 150 python bench.py $S | line
echo "-- defaults again";                             timeout 150 python bench.py $S | line
echo "== done"
