"""Pipeline parallelism for a model the library has never seen: the graph partitioner traces it with torch.fx, cuts it
where exactly one activation is live, and the static 1F1B engine trains it.

    torchrun --standalone --nnodes=1 --nproc-per-node 4 examples/pipeline_custom_model.py --pp 2 --dp 2 --backend gloo

The model below is a small pre-norm transformer over byte tokens with a padding mask; nothing in it follows a 🤗 naming
convention.  ``mask`` is an *input-derived* value: it never travels between the stages, every stage reads it from its
copy of the micro-batch.  The task (copy the previous token) is learnable in a few dozen steps.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch
from torch import nn

from pipegoose_b200.distributed import ParallelContext, ParallelMode
from pipegoose_b200.nn import DataParallel, PipelineParallel
from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType, get_scheduler
from pipegoose_b200.optim import DistributedOptimizer


class Block(nn.Module):
    def __init__(self, d: int, heads: int):
        super().__init__()
        self.heads = heads
        self.ln1, self.ln2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.qkv, self.proj = nn.Linear(d, 3 * d), nn.Linear(d, d)
        self.up, self.down = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

    def forward(self, x, bias):
        b, s, d = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(b, s, 3, self.heads, d // self.heads).permute(2, 0, 3, 1, 4)
        att = torch.softmax(q @ k.transpose(-1, -2) / (d // self.heads) ** 0.5 + bias, dim=-1)
        x = x + self.proj((att @ v).transpose(1, 2).reshape(b, s, d))
        return x + self.down(torch.nn.functional.gelu(self.up(self.ln2(x))))


class ByteLM(nn.Module):
    def __init__(self, vocab=64, d=64, heads=4, layers=4, max_len=64):
        super().__init__()
        self.tok, self.pos = nn.Embedding(vocab, d), nn.Embedding(max_len, d)
        self.blocks = nn.ModuleList([Block(d, heads) for _ in range(layers)])
        self.norm, self.out = nn.LayerNorm(d), nn.Linear(d, vocab)
        self.register_buffer("causal", torch.full((max_len, max_len), float("-inf")).triu(1))
        self.register_buffer("positions", torch.arange(max_len))
        self.seq_len, self.vocab = max_len, vocab

    def forward(self, tokens, mask, labels):
        # [batch, 1, seq, seq] additive attention bias from the causal pattern and the padding mask: input-derived
        bias = self.causal[: self.seq_len, : self.seq_len] + (1.0 - mask.float())[:, None, None, :] * -1e9
        x = self.tok(tokens) + self.pos(self.positions[: self.seq_len])
        for blk in self.blocks:
            x = blk(x, bias)
        logits = self.out(self.norm(x))
        return nn.functional.cross_entropy(logits.view(-1, self.vocab), labels.view(-1), ignore_index=-100)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pp", type=int, default=2)
    ap.add_argument("--dp", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--microbatches", type=int, default=4)
    ap.add_argument("--backend", default="nccl" if torch.cuda.is_available() else "gloo")
    args = ap.parse_args()

    ctx = ParallelContext.from_torch(tensor_parallel_size=1, pipeline_parallel_size=args.pp, data_parallel_size=args.dp,
                                     backend=args.backend)
    seq = 16
    torch.manual_seed(0)
    model = ByteLM(max_len=seq)
    model = PipelineParallel(model, num_microbatches=args.microbatches, parallel_context=ctx,
                             scheduler_type=SchedulerType.ONE_F_ONE_B).parallelize()
    if args.dp > 1:
        model = DataParallel(model, ctx).parallelize()
    if args.backend == "nccl":
        model.to("cuda")
    stage = model._pg_pipeline_stage
    optim = torch.optim.Adam([p for p in stage.parameters()], lr=3e-3)
    if args.dp > 1:
        optim = DistributedOptimizer(optim, ctx)

    sched = get_scheduler(SchedulerType.ONE_F_ONE_B)(args.microbatches, args.pp)
    if ctx.get_global_rank() == 0:
        print(f"stage 0 of {args.pp}: {type(stage).__name__} reading {stage.stage_inputs} from the micro-batch; "
              f"schedule bubble {sched.bubble_fraction():.1%}", flush=True)

    g = torch.Generator().manual_seed(100 + ctx.get_local_rank(ParallelMode.DATA))
    first = last = None
    for step in range(args.steps):
        tokens = torch.randint(1, 64, (8, seq), generator=g)
        lengths = torch.randint(seq // 2, seq + 1, (8, 1), generator=g)
        mask = (torch.arange(seq)[None, :] < lengths).long()
        labels = torch.roll(tokens, 1, dims=1)                 # predict the previous token
        labels[:, 0] = -100
        labels = labels.masked_fill(mask == 0, -100)
        out = model(tokens, mask=mask, labels=labels)
        optim.zero_grad()
        out.loss.backward()
        optim.step()
        last = float(out.loss.detach())
        first = last if first is None else first
        if ctx.get_global_rank() == 0 and step % 10 == 0:
            print(f"step {step:3d}  loss {last:.4f}", flush=True)
    if ctx.get_global_rank() == 0:
        print(f"loss {first:.4f} -> {last:.4f}", flush=True)
    assert last < first, "the loss did not go down"
    ctx.destroy()


if __name__ == "__main__":
    main()
