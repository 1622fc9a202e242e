"""3D-parallel Bloom training with pipegoose_b200 (the reference's examples/hybrid_parallelism.py flow).

    torchrun --standalone --nnodes=1 --nproc-per-node 4 examples/hybrid_parallelism.py --tp 2 --dp 2

Synthetic token ids are used (no dataset / tokenizer download needed).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch

from pipegoose_b200.distributed import ParallelContext, ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--model", default="bloom_560m", help="bloom_tiny | bloom_560m | bloom_1b7 | bloom_3b | bloom_7b1 | gpt2_tiny | gpt2 | gpt2_medium | gpt2_large")
    ap.add_argument("--batch", type=int, default=8, help="sequences per data-parallel replica")
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--microbatches", type=int, default=4)
    ap.add_argument("--backend", default="nccl" if torch.cuda.is_available() else "gloo")
    ap.add_argument("--hf", action="store_true",
                    help="start from a transformers BloomForCausalLM, exactly like the reference's README: TensorParallel "
                         "converts it in place to the fused sequence-parallel path")
    args = ap.parse_args()

    parallel_context = ParallelContext.from_torch(
        tensor_parallel_size=args.tp, pipeline_parallel_size=args.pp, data_parallel_size=args.dp, backend=args.backend)
    rank = parallel_context.get_global_rank()

    if args.model.startswith("gpt2"):  # the GPT-2 family runs on the same fused blocks (models/gpt2.py)
        from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel

        cfg = getattr(GPT2Config, args.model)()
        model = GPT2LMHeadModel(cfg)
    elif args.hf:
        from transformers import BloomConfig as HFBloomConfig
        from transformers import BloomForCausalLM as HFBloomForCausalLM

        cfg = getattr(BloomConfig, args.model)()
        model = HFBloomForCausalLM(HFBloomConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, n_layer=cfg.n_layer,
                                                 n_head=cfg.n_head))
    else:
        cfg = getattr(BloomConfig, args.model)()
        model = BloomForCausalLM(cfg)
    if args.backend == "nccl":
        model = model.to(torch.bfloat16)
    # (bf16 transformers models take the fused path by default; sequence_parallel=True asks for it for fp32 CPU runs too)
    model = TensorParallel(model, parallel_context, sequence_parallel=True if args.hf else None).parallelize()
    if args.pp > 1:
        model = PipelineParallel(model, num_microbatches=args.microbatches, parallel_context=parallel_context).parallelize()
    model = DataParallel(model, parallel_context).parallelize()
    if args.backend == "nccl":
        model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-4), parallel_context)

    device = next(model.parameters()).device
    gen = torch.Generator().manual_seed(parallel_context.get_local_rank(ParallelMode.DATA))
    for step in range(args.steps):
        ids = torch.randint(0, cfg.vocab_size, (args.batch, args.seq), generator=gen).to(device)
        outputs = model(ids, labels=ids)
        optim.zero_grad()
        outputs.loss.backward()
        optim.step()
        if rank == parallel_context.get_world_size(ParallelMode.GLOBAL) - 1 or args.pp == 1 and rank == 0:
            print(f"step {step} loss {outputs.loss.item():.4f}", flush=True)
    parallel_context.destroy()
