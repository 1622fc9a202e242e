#!/usr/bin/env python
"""Flagship benchmark: bloom-560m training tokens/s (TP2 x DP, weak scaling) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3
    python bench.py --impl reference ...      # the unmodified reference (baseline/_ref) on the same config

Metric/config follow BASELINE.json: bloom-560m, seq 1024, bf16, synthetic token ids, random-init
weights, Adam; N=1 -> TP1xDP1, N>=2 -> TP2 x DP(N/2); per-GPU work is fixed (weak scaling):
global batch = batch_per_gpu * N sequences.  One JSON line is printed by rank 0.

Two timed regions of K steps each, both bracketed by barrier + synchronize, timed with CUDA events
on the device, max over ranks:
  value : full train step (fwd, bwd, grad sync, optimizer) with device-resident inputs
  e2e   : the same step through the public API including, every step, the pinned-host -> device
          copy of that step's token ids and a device -> host read of the loss.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _env_defaults():
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "MASTER_PORT" not in os.environ:
        os.environ["MASTER_PORT"] = str(_free_port())


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.samples.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[1]))
                mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def parallel_layout(n_gpus: int):
    if n_gpus == 1:
        return 1, 1
    return 2, n_gpus // 2


def timed_region(torch, dist, steps, fn):
    """barrier + sync, K steps between two CUDA events, sync + barrier; returns max-over-ranks ms."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        fn(i)
    end.record()
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(end)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item())


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from pipegoose_b200 import ops
    from pipegoose_b200.distributed import ParallelContext, ParallelMode
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer
    from pipegoose_b200.optim.fused_adam import FusedAdam

    tp, dp = parallel_layout(args.gpus)
    pp = max(args.pp, 1)
    if args.tp > 0 or pp > 1:
        # the other BASELINE.json configs: --tp 8 (bloom-7b1), --tp 2 --pp 2 (bloom-3b, 1F1B), --experts 8 (Switch MoE)
        tp = args.tp if args.tp > 0 else 1
        assert args.gpus % (tp * pp) == 0, "--gpus must be a multiple of tp * pp"
        dp = args.gpus // (tp * pp)
    ctx = ParallelContext.from_torch(tensor_parallel_size=tp, pipeline_parallel_size=pp, data_parallel_size=dp,
                                     backend="nccl")
    rank = ctx.get_global_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(1234)
    if args.model.startswith("gpt2"):  # not a BASELINE.json config: the second model family on the same kernels
        from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel

        cfg = getattr(GPT2Config, args.model.replace("-", "_"))()
        model = GPT2LMHeadModel(cfg).to(torch.bfloat16)
    else:
        cfg = getattr(BloomConfig, args.model.replace("-", "_"))()
        model = BloomForCausalLM(cfg).to(torch.bfloat16)
    if args.experts > 0:
        from pipegoose_b200.nn import ExpertParallel
        from pipegoose_b200.nn.expert_parallel import SwitchNoisePolicy, Top1Router

        router = Top1Router(SwitchNoisePolicy(), args.experts, cfg.hidden_size, expert_capacity=(1.25, 2.0))
        layers = list(range(0, cfg.n_layer, max(args.moe_every, 1)))
        model = ExpertParallel(model, args.experts, mapping=layers, router=router.to(torch.bfloat16),
                               parallel_context=ctx).parallelize()
    model = TensorParallel(model, ctx).parallelize()
    if pp > 1:
        from pipegoose_b200.nn import PipelineParallel

        model = PipelineParallel(model, num_microbatches=args.microbatches, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=args.lr), ctx)

    S = args.seq_len
    b_rep = args.batch_per_gpu * tp * pp  # sequences per model replica (DP rank)
    gen = torch.Generator().manual_seed(1000 + ctx.get_local_rank(ParallelMode.DATA))
    n_host = args.steps + args.warmup + 1
    host_ids = [torch.randint(0, cfg.vocab_size, (b_rep, S), generator=gen).pin_memory() for _ in range(n_host)]
    dev_ids = host_ids[0].to(dev)

    def total_loss(ids):
        loss = model(ids, labels=ids).loss
        if args.experts > 0:
            # Switch training objective: task loss + load-balancing and router-z terms the MoE layers pushed
            from pipegoose_b200.nn.expert_parallel import ExpertContext

            store = ExpertContext.get_instance()
            loss = loss + 0.01 * sum(store.pop_all_aux_loss()) + 0.001 * sum(store.pop_all_z_loss())
        return loss

    def step_device(i):
        loss = total_loss(dev_ids)
        optim.zero_grad()
        loss.backward()
        optim.step()
        return loss

    last = {}

    def step_e2e(i):
        ids = host_ids[i % n_host].to(dev, non_blocking=True)
        loss = total_loss(ids)
        optim.zero_grad()
        loss.backward()
        optim.step()
        last["loss"] = loss.item()  # device -> host read of the step's result (0 on non-final pipeline stages)

    for i in range(args.warmup):
        step_e2e(i)
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    ms_dev = timed_region(torch, dist, args.steps, step_device)
    launches = ops.launch_count()
    ms_e2e = timed_region(torch, dist, args.steps, step_e2e)
    if rank == 0:
        sampler.stop()
    tokens_per_step = args.batch_per_gpu * args.gpus * S
    result = {
        "metric": "bloom-560m training tokens/sec (whole job, device-timed, max over ranks)" if args.model == "bloom-560m"
        else f"{args.model} training tokens/sec (whole job, device-timed, max over ranks)",
        "value": tokens_per_step * args.steps / (ms_dev / 1e3),
        "unit": "tokens/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic token ids (uniform random), random-init weights",
        "impl": "pipegoose_b200",
        "config": {
            "model": args.model, "global_batch": args.batch_per_gpu * args.gpus, "seq_len": S,
            "parallelism": f"tp{tp}" + (f"pp{pp}(1f1b,{args.microbatches}mb)" if pp > 1 else "") + f"dp{dp}"
            + ("+zero1" if dp > 1 else "") + (f"+moe{args.experts}e" if args.experts > 0 else ""),
            "optimizer": "Adam (fp32 master + moments, fused)", "batch_per_gpu": args.batch_per_gpu,
            "l2": "no explicit flush: per-step working set (1.1 GB bf16 weights + >6 GB activations) >> 126 MB L2",
        },
        "e2e": {
            "value": tokens_per_step * args.steps / (ms_e2e / 1e3), "unit": "tokens/s",
            "ms_per_step": ms_e2e / args.steps,
            "h2d_bytes_per_step": int(host_ids[0].numel() * host_ids[0].element_size()),
            "d2h_bytes_per_step": 4,
        },
        "gpu_launches": launches,
        "final_loss": last.get("loss"),
        "mfu_vs_measured_sustained": None,
        "clocks": sampler.summary() if rank == 0 else None,
    }
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        flops = model.flops_per_token(S) * tokens_per_step / args.gpus
        result["mfu_vs_measured_sustained"] = flops / (ms_dev / args.steps / 1e3) / (peaks["bf16_tflops_sustained"] * 1e12)
    except Exception:
        pass
    if rank == 0:
        print(json.dumps(result), flush=True)
    ctx.destroy()


# --------------------------------------------------------------------------------------------
# reference arm: the unmodified reference installed under baseline/_ref, stock code path
# --------------------------------------------------------------------------------------------
def run_reference(args):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "pipegoose")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/pipegoose not installed (see DESIGN.md)"}))
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
    sys.path.insert(0, ref_dir)
    try:
        import hf_fx_shim

        hf_fx_shim.install()
        import torch
        import torch.distributed as dist
        from pipegoose.distributed.parallel_context import ParallelContext
        from pipegoose.distributed.parallel_mode import ParallelMode
        from pipegoose.nn import DataParallel, TensorParallel
        from pipegoose.optim import DistributedOptimizer
        from transformers import BloomConfig, BloomForCausalLM
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": f"import failed: {type(e).__name__}: {e}"[:300]}))
        return

    tp, dp = parallel_layout(args.gpus)
    # The reference issues collectives on CPU tensors during bring-up (parallel_context.py:263-287), which a
    # pure "nccl" group rejects; torch's per-device backend string gives it gloo for those and NCCL for CUDA
    # tensors.  This is an argument of the reference's public API, not a change to it.
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    ctx = ParallelContext.from_torch(tensor_parallel_size=tp, pipeline_parallel_size=1, data_parallel_size=dp,
                                     backend="cpu:gloo,cuda:nccl")
    ctx.set_device()
    rank = ctx.get_global_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = {"bloom-560m": (1024, 24, 16), "bloom-3b": (2560, 30, 32), "bloom-7b1": (4096, 30, 32), "bloom-1b7": (2048, 24, 16)}
    h, L, nh = sizes[args.model]
    torch.manual_seed(1234)
    cfg = BloomConfig(hidden_size=h, n_layer=L, n_head=nh, vocab_size=250880)
    model = BloomForCausalLM(cfg).to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = torch.optim.Adam(model.parameters(), lr=args.lr)
    optim = DistributedOptimizer(optim, ctx)
    model.to("cuda")
    model.train()

    S = args.seq_len
    b_rep = args.batch_per_gpu * tp
    gen = torch.Generator().manual_seed(1000 + ctx.get_local_rank(ParallelMode.DATA))
    n_host = args.steps + args.warmup + 1
    host_ids = [torch.randint(0, cfg.vocab_size, (b_rep, S), generator=gen).pin_memory() for _ in range(n_host)]
    dev_ids = host_ids[0].to(dev)
    mask = torch.ones(b_rep, S, dtype=torch.long, device=dev)

    def step_device(i):
        out = model(input_ids=dev_ids, attention_mask=mask, labels=dev_ids)
        optim.zero_grad()
        out.loss.backward()
        optim.step()

    last = {}

    def step_e2e(i):
        ids = host_ids[i % n_host].to(dev, non_blocking=True)
        out = model(input_ids=ids, attention_mask=mask, labels=ids)
        optim.zero_grad()
        out.loss.backward()
        optim.step()
        last["loss"] = out.loss.item()

    for i in range(args.warmup):
        step_e2e(i)
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    ms_dev = timed_region(torch, dist, args.steps, step_device)
    ms_e2e = timed_region(torch, dist, args.steps, step_e2e)
    if rank == 0:
        sampler.stop()
    tokens_per_step = args.batch_per_gpu * args.gpus * S
    result = {
        "impl": "reference",
        "metric": f"{args.model} training tokens/sec (whole job, device-timed, max over ranks)",
        "value": tokens_per_step * args.steps / (ms_dev / 1e3), "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic token ids (uniform random), random-init weights",
        "config": {"model": args.model, "global_batch": args.batch_per_gpu * args.gpus, "seq_len": S,
                   "parallelism": f"tp{tp}dp{dp}" + ("+zero1" if dp > 1 else ""),
                   "optimizer": "torch.optim.Adam via reference DistributedOptimizer",
                   "batch_per_gpu": args.batch_per_gpu},
        "e2e": {"value": tokens_per_step * args.steps / (ms_e2e / 1e3), "unit": "tokens/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(host_ids[0].numel() * host_ids[0].element_size()), "d2h_bytes_per_step": 4},
        "gpu_launches": 0,
        "final_loss": last.get("loss"),
        "clocks": sampler.summary() if rank == 0 else None,
    }
    if rank == 0:
        print(json.dumps(result), flush=True)
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="bloom-560m")
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--batch-per-gpu", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-4)
    # non-default layouts (ours arm only): BASELINE.json configs #3-#5
    ap.add_argument("--tp", type=int, default=0, help="tensor parallel size (0: 1 GPU -> 1, else 2)")
    ap.add_argument("--pp", type=int, default=1, help="pipeline stages (1F1B)")
    ap.add_argument("--microbatches", type=int, default=8)
    ap.add_argument("--experts", type=int, default=0, help="Switch-MoE experts (sharded over the tensor group)")
    ap.add_argument("--moe-every", type=int, default=2, help="every n-th block gets a MoE MLP")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    _env_defaults()
    world = int(os.environ["WORLD_SIZE"])
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.abspath(__file__)] + sys.argv[1:]
            os.execv(sys.executable, cmd)
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    if args.impl == "reference":
        try:
            run_reference(args)
        except Exception as e:  # the reference's own code path failed on this layout: report it, do not crash the driver
            import traceback

            traceback.print_exc()
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "n_gpus": args.gpus,
                                  "unavailable": f"reference failed at run time: {type(e).__name__}: {e}"[:300]}), flush=True)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
