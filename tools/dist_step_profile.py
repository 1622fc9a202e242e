"""Per-kernel breakdown of one multi-GPU training step (diagnosis only — never a bench number).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_step_profile.py [--tp 2] [--zero 1]

Rank 0 records one step with torch.profiler (CUPTI) after warm-up and writes the per-kernel device
time table to gpurun_out/dist_profile_tp{tp}dp{dp}.json; every rank also reports event-timed
fwd / bwd / optimizer phases (max over ranks).
"""
import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--model", default="bloom_560m")
    ap.add_argument("--batch-per-gpu", type=int, default=8)
    ap.add_argument("--seq-len", type=int, default=1024)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile

    from pipegoose_b200.distributed import ParallelContext
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer
    from pipegoose_b200.optim.fused_adam import FusedAdam

    world = int(os.environ["WORLD_SIZE"])
    tp = args.tp
    dp = world // tp
    ctx = ParallelContext.from_torch(tensor_parallel_size=tp, pipeline_parallel_size=1, data_parallel_size=dp,
                                     backend="nccl")
    rank = ctx.get_global_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = getattr(BloomConfig, args.model)()
    torch.manual_seed(1234)
    model = BloomForCausalLM(cfg).to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-4), ctx)
    ids = torch.randint(0, cfg.vocab_size, (args.batch_per_gpu * tp, args.seq_len), device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        if timed:
            ev[0].record()
        loss = model(ids, labels=ids).loss
        if timed:
            ev[1].record()
        optim.zero_grad()
        loss.backward()
        if timed:
            ev[2].record()
        optim.step()
        if timed:
            ev[3].record()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    step(timed=True)
    torch.cuda.synchronize()
    phases = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], device=dev)
    dist.all_reduce(phases, op=dist.ReduceOp.MAX)
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        agg = defaultdict(lambda: [0, 0.0])
        for e in prof.events():
            if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time > 0:
                agg[e.name.split("(")[0]][0] += 1
                agg[e.name.split("(")[0]][1] += e.device_time
        total = sum(v[1] for v in agg.values())
        table = sorted(([k, v[0], v[1]] for k, v in agg.items()), key=lambda t: -t[2])
        out = {"tp": tp, "dp": dp, "phases_ms_max_over_ranks": {"fwd": phases[0].item(), "bwd": phases[1].item(),
                                                                 "optim": phases[2].item()},
               "sum_kernel_us": total,
               "kernels": [{"name": k, "count": c, "us": round(us, 1), "frac": round(us / total, 4)} for k, c, us in table]}
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/dist_profile_tp{tp}dp{dp}.json", "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out["phases_ms_max_over_ranks"]))
        for k in out["kernels"][:30]:
            print(f"{k['us']:10.1f} us  {k['frac']*100:5.1f}%  x{k['count']:<4d} {k['name'][:90]}")
        print(f"sum of kernel time {total/1e3:.2f} ms")
    ctx.destroy()


if __name__ == "__main__":
    main()
