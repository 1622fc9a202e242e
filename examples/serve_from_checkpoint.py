"""From a sharded training checkpoint to generation on another layout.

Phase 1 (``--phase train``, e.g. 4 ranks, TP2 x DP2 + ZeRO-1): a tiny Bloom learns arithmetic progressions and writes
its (tp, pp) shards with ``save_pretrained``.
Phase 2 (``--phase serve``, ANY tensor-parallel size, here 1): the shards are merged offline
(``nn.checkpoint_convert.consolidate_checkpoint``), loaded into a fresh model, parallelized for the serving layout and
``generate()`` continues a batch of LEFT-padded prompts of unequal length with a KV cache (greedy and nucleus sampling).

    torchrun --standalone --nproc-per-node 4 examples/serve_from_checkpoint.py --phase train --tp 2 --dp 2 --backend gloo --ckpt /tmp/pg_serve
    python examples/serve_from_checkpoint.py --phase serve --train-tp 2 --ckpt /tmp/pg_serve
    torchrun --standalone --nproc-per-node 2 examples/serve_from_checkpoint.py --phase serve --train-tp 2 --tp 2 --backend gloo --ckpt /tmp/pg_serve
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch

from pipegoose_b200.distributed import ParallelContext, ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.nn.checkpoint_convert import consolidate_checkpoint
from pipegoose_b200.nn.utils import save_pretrained
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam

CFG = dict(vocab_size=250, hidden_size=128, n_layer=2, n_head=4)     # 250: not a multiple of 8 x tp (padding is exercised)
STEP = 3


def train(args, ctx):
    torch.manual_seed(0)
    cfg = BloomConfig(**CFG)
    model = BloomForCausalLM(cfg)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=3e-3), ctx)
    g = torch.Generator().manual_seed(1 + ctx.get_local_rank(ParallelMode.DATA))
    for step in range(args.steps):
        start = torch.randint(0, cfg.vocab_size, (8, 1), generator=g)
        ids = (start + STEP * torch.arange(32)[None, :]) % cfg.vocab_size
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        if ctx.get_global_rank() == 0 and (step % 50 == 0 or step == args.steps - 1):
            print(f"step {step} loss {loss.item():.4f}", flush=True)
    save_pretrained(model, ckp_path=args.ckpt, parallel_context=ctx)
    torch.distributed.barrier()             # every (tp, pp) writer is done
    if ctx.get_global_rank() == 0:
        print(f"wrote {sorted(os.listdir(args.ckpt))}", flush=True)


def serve(args, ctx):
    cfg = BloomConfig(**CFG)
    state = consolidate_checkpoint(args.ckpt, tensor_parallel_size=args.train_tp, pipeline_parallel_size=1)
    model = BloomForCausalLM(cfg)
    model.load_state_dict(state)            # strict: the merged dict IS the unsharded model
    model = TensorParallel(model, ctx).parallelize().eval()
    # three prompts of 6, 4 and 2 tokens, left-padded to one batch (what a 🤗 tokenizer with padding_side="left" returns)
    lengths, width = (6, 4, 2), 6
    ids = torch.zeros(3, width, dtype=torch.long)
    mask = torch.zeros(3, width, dtype=torch.long)
    for r, (n, start) in enumerate(zip(lengths, (5, 100, 200))):
        ids[r, width - n:] = (start + STEP * torch.arange(n)) % cfg.vocab_size
        mask[r, width - n:] = 1
    out = model.generate(ids, attention_mask=mask, max_new_tokens=8)
    sampled = model.generate(ids, attention_mask=mask, max_new_tokens=8, do_sample=True, top_p=0.9, temperature=0.7,
                             generator=torch.Generator().manual_seed(0))
    if ctx.get_global_rank() == 0:
        right = 0
        for r, n in enumerate(lengths):
            want = (ids[r, -1] + STEP * torch.arange(1, 9)) % cfg.vocab_size
            right += int((out[r, width:] == want).sum())
            print(f"prompt {ids[r, width - n:].tolist()} -> greedy {out[r, width:].tolist()}  nucleus {sampled[r, width:].tolist()}",
                  flush=True)
        print(f"served at tp={ctx.tensor_parallel_size} from a tp={args.train_tp} checkpoint: {right}/24 greedy tokens continue "
              "the progression", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", choices=("train", "serve"), required=True)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--dp", type=int, default=1)
    ap.add_argument("--train-tp", type=int, default=2, help="serve: tensor-parallel size the checkpoint was written with")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--ckpt", default="/tmp/pg_serve")
    ap.add_argument("--backend", default="gloo")
    args = ap.parse_args()
    os.environ.setdefault("RANK", "0"), os.environ.setdefault("WORLD_SIZE", "1"), os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"), os.environ.setdefault("MASTER_PORT", "29533")
    ctx = ParallelContext.from_torch(tensor_parallel_size=args.tp, pipeline_parallel_size=1, data_parallel_size=args.dp,
                                     backend=args.backend)
    (train if args.phase == "train" else serve)(args, ctx)
    ctx.destroy()


if __name__ == "__main__":
    main()
