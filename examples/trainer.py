"""The Trainer: gradient accumulation, global-norm clipping, LR schedule, periodic sharded checkpoints, resume, watchdog.

    torchrun --standalone --nnodes=1 --nproc-per-node 4 examples/trainer.py --tp 2 --dp 2 --backend gloo --ckpt /tmp/pg_ckpt
    (run it twice: the second run resumes from the last checkpoint of the first)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch

from pipegoose_b200.distributed import ParallelContext, ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.trainer import DistributedLogger, Trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20, help="optimizer steps per run")
    ap.add_argument("--ckpt", default="")
    ap.add_argument("--backend", default="nccl" if torch.cuda.is_available() else "gloo")
    args = ap.parse_args()

    ctx = ParallelContext.from_torch(tensor_parallel_size=args.tp, pipeline_parallel_size=1, data_parallel_size=args.dp,
                                     backend=args.backend)
    cfg = BloomConfig(vocab_size=256, hidden_size=128, n_layer=2, n_head=4)
    torch.manual_seed(0)
    model = BloomForCausalLM(cfg)
    gpu = args.backend == "nccl"
    if gpu:
        model = model.to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    if gpu:
        model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=3e-3, weight_decay=0.01, adamw=True), ctx)
    warmup = 5
    sched = torch.optim.lr_scheduler.LambdaLR(optim.optim, lambda s: min(1.0, (s + 1) / warmup))

    accum = 2
    g = torch.Generator().manual_seed(100 + ctx.get_local_rank(ParallelMode.DATA))
    start = torch.randint(0, cfg.vocab_size, (args.steps * accum * 2, 8, 1), generator=g)
    data = [{"input_ids": (s + 3 * torch.arange(32)[None, :]) % cfg.vocab_size} for s in start]   # 2 runs' worth of batches

    trainer = Trainer(model, data, optim=optim, parallel_context=ctx, grad_accum_steps=accum, max_grad_norm=1.0,
                      lr_scheduler=sched, log_every=5, loggers=[DistributedLogger(ctx)], checkpoint_dir=args.ckpt or None,
                      checkpoint_every=10 if args.ckpt else 0, resume=bool(args.ckpt), watchdog_timeout_s=120)
    resumed_at = trainer.state.step if not trainer.load_checkpoint() else trainer.state.step
    trainer.resume = False                                  # (already loaded above, to know where this run starts)
    trainer.max_steps = resumed_at + args.steps             # the loader is positional: resume skips what was consumed
    state = trainer.fit()
    if ctx.get_global_rank() == 0:
        print(f"optimizer steps {resumed_at} -> {state.step}, last loss {state.last_loss:.4f}, "
              f"last grad norm {state.last_grad_norm:.3f}", flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()
