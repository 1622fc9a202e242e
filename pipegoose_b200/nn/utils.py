"""Checkpoint I/O (parity: reference nn/utils.py:11-50): one file per (tp_rank, pp_rank) named
``pytorch_model_tp_{tp}_pp_{pp}.bin``.  Only data-parallel rank 0 writes (the reference has every
replica write the same path) and the directory is created when missing."""
from __future__ import annotations

import os
from pathlib import Path

import torch
from torch import nn

from pipegoose_b200.constants import CHECKPOINT_PATH_NAME, CHECKPOINT_WEIGHTS_NAME
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def _ckpt_file(ckp_path: str, ckp_name: str, parallel_context: ParallelContext) -> str:
    tp_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    pp_rank = parallel_context.get_local_rank(ParallelMode.PIPELINE)
    return os.path.join(ckp_path, ckp_name.format(tp_rank, pp_rank))


def from_pretrained(module: nn.Module, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                    ckp_name: str = CHECKPOINT_WEIGHTS_NAME):
    """Load this rank's (tp, pp) shard into an already parallelized ``module``."""
    path = _ckpt_file(ckp_path, ckp_name, parallel_context)
    if not os.path.exists(path):
        raise ValueError(f"ckp_path {path} does not exist")
    state_dict = torch.load(path, map_location="cpu")
    with torch.no_grad():
        own = module.state_dict()
        for k, v in state_dict.items():
            if k in own:
                own[k].copy_(v)
            else:
                raise KeyError(f"unexpected key {k} in checkpoint {path}")
    return module


def save_pretrained(module: nn.Module, ckp_name: str = CHECKPOINT_WEIGHTS_NAME, ckp_path: str = CHECKPOINT_PATH_NAME,
                    parallel_context: ParallelContext = None):
    """Save this rank's (tp, pp) shard; data-parallel replicas > 0 skip the write."""
    Path(ckp_path).mkdir(parents=True, exist_ok=True)
    if parallel_context.get_local_rank(ParallelMode.DATA) == 0:
        state = {k: v.detach().cpu() for k, v in module.state_dict().items()}
        torch.save(state, _ckpt_file(ckp_path, ckp_name, parallel_context))
    if parallel_context.get_world_size(ParallelMode.DATA) > 1:
        import torch.distributed as dist

        dist.barrier(group=parallel_context.get_group(ParallelMode.DATA))
