#!/usr/bin/env bash
# 1-GPU A/B of the kernel variants built by tools/build_variants.sh (one gpurun call, ~6 min):
#   gpurun --timeout 900 -- 'bash tools/ab_1gpu.sh > gpurun_out/ab_1gpu.log 2>&1'
set -uo pipefail
mkdir -p gpurun_out
bench_line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ', round(d['ms_per_step'], 2), 'ms/step  e2e', round(d['e2e']['ms_per_step'], 2), ' loss', d['final_loss'], d['clocks'])
"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "-- main";                   python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
if [ -f pipegoose_b200/_C_pdl.so ]; then
  echo "-- pdl: kernel tests (PDL=1)"; PIPEGOOSE_B200_EXT=pdl PIPEGOOSE_B200_PDL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
  echo "-- pdl (attribute off)";  PIPEGOOSE_B200_EXT=pdl python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
  echo "-- pdl (PDL=1)";          PIPEGOOSE_B200_EXT=pdl PIPEGOOSE_B200_PDL=1 python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
fi
if [ -f pipegoose_b200/_C_coresident.so ]; then
  echo "-- coresident: kernel tests"; PIPEGOOSE_B200_EXT=coresident timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
  echo "-- coresident 1 GPU (cost of 152 registers)"; PIPEGOOSE_B200_EXT=coresident python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
fi
echo "-- main again";             python bench.py --gpus 1 --steps 10 --warmup 3 | bench_line
echo "-- torch symmetric memory / multicast support probe"
python - <<'EOF'
import os, torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("nccl", rank=0, world_size=1)
torch.cuda.set_device(0)
import torch.distributed._symmetric_memory as sm
try:
    t = sm.empty(1 << 20, dtype=torch.uint8, device="cuda")
    h = sm.rendezvous(t, dist.group.WORLD)
    print("symm ok: has_multicast_support", getattr(h, "has_multicast_support", None), "mc_ptr", hex(h.multicast_ptr), "bufs", [hex(p) for p in h.buffer_ptrs], "sigpad", h.signal_pad_size)
except Exception as e:
    print("symm failed:", type(e).__name__, e)
try:
    from torch._C._distributed_c10d import _SymmetricMemory
    print("has_multicast_support(cuda,0):", _SymmetricMemory.has_multicast_support(torch.device("cuda").type if False else "cuda", 0))
except Exception as e:
    print("has_multicast_support probe failed:", type(e).__name__, e)
EOF
