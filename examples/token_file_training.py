"""Training from a pre-tokenised binary file: TokenFileDataset (memmap) -> build_dataloader (one shard per data-parallel
replica, the same batches inside a tensor group, pinned memory + device prefetch on GPUs) -> Trainer.

    torchrun --standalone --nnodes=1 --nproc-per-node 4 examples/token_file_training.py --tp 2 --dp 2 --backend gloo
"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch

from pipegoose_b200.distributed import ParallelContext
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.trainer import DistributedLogger, Trainer
from pipegoose_b200.utils.data import TokenFileDataset, build_dataloader, write_token_file


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--tokens", default="", help="flat uint16 token file (default: a synthetic one is written)")
    ap.add_argument("--backend", default="nccl" if torch.cuda.is_available() else "gloo")
    args = ap.parse_args()

    ctx = ParallelContext.from_torch(tensor_parallel_size=args.tp, pipeline_parallel_size=1, data_parallel_size=args.dp,
                                     backend=args.backend)
    cfg = BloomConfig(vocab_size=256, hidden_size=128, n_layer=2, n_head=4)
    path = args.tokens
    if not path:
        # a learnable synthetic corpus: arithmetic progressions modulo the vocabulary (rank 0 writes, everybody reads)
        path = os.path.join(tempfile.gettempdir(), "pipegoose_b200_example_tokens.bin")
        if ctx.get_global_rank() == 0:
            g = torch.Generator().manual_seed(0)
            starts = torch.randint(0, cfg.vocab_size, (512, 1), generator=g)
            write_token_file(path, ((starts + 3 * torch.arange(32)[None, :]) % cfg.vocab_size).reshape(-1))
        torch.distributed.barrier()
    dataset = TokenFileDataset(path, seq_len=32)
    loader = build_dataloader(dataset, ctx, batch_size=8, shuffle=True)

    torch.manual_seed(0)
    model = BloomForCausalLM(cfg)
    gpu = args.backend == "nccl"
    if gpu:
        model = model.to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    if gpu:
        model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=3e-3), ctx)
    trainer = Trainer(model, loader, optim=optim, parallel_context=ctx, num_epochs=args.epochs, max_grad_norm=1.0,
                      log_every=16, loggers=[DistributedLogger(ctx)])
    state = trainer.fit()
    if ctx.get_global_rank() == 0:
        print(f"{len(dataset)} sequences, {len(loader)} batches per replica and epoch, {state.step} steps, "
              f"last loss {state.last_loss:.4f}", flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()
