"""Base class of the in-place parallel wrappers (parity: reference nn/parallel.py:19-93).

``X(module, ..., parallel_context).parallelize()`` mutates and returns the same ``nn.Module``.
After a wrapper with group size > 1 ran, the module carries ``parallel_metadata`` and its
``.to()`` / ``.cuda()`` place it on this rank's GPU.  Differences from the reference: the local
device is the process's local rank (not ``device % world_size``), ``.to`` accepts ``"cuda"``,
``"cuda:N"``, a ``torch.device`` or a dtype and returns the module, and flat parameter buffers
are re-bound after the move.
"""
from __future__ import annotations

from abc import abstractmethod
from dataclasses import dataclass
from functools import partial
from typing import Optional, Union

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


@dataclass
class ParallelMetadata:
    device: Optional[int] = None
    local_device: Optional[int] = None
    is_sliced: bool = False
    # how a sliced parameter was cut (checkpoint consolidation / resharding, nn/checkpoint_convert.py): the dimension
    # split over the tensor group and that dimension's size in the unsharded model (before any zero padding)
    partition_dim: Optional[int] = None
    full_size: Optional[int] = None
    is_vocab: bool = False      # a vocabulary table / lm_head: zero-padded to a multiple of the group before the cut
    vocab_multiple: int = 1     # ... of ``vocab_multiple x group size`` (8 on the sequence-parallel fast path: kernel tiles)


class Parallel:
    """Base of DataParallel / TensorParallel / PipelineParallel / ExpertParallel."""

    def __init__(self, module: nn.Module, parallel_context: ParallelContext):
        self.module = module
        self.parallel_context = parallel_context

    @abstractmethod
    def parallelize(self) -> nn.Module:
        raise NotImplementedError

    @abstractmethod
    def deparallelize(self) -> nn.Module:
        raise NotImplementedError

    def _save_metadata(self, module: nn.Module, parallel_context: ParallelContext):
        ctx = parallel_context

        def local(mode):
            return ctx.get_local_rank(mode) if ctx.is_initialized(mode) else 0

        device = ctx.ranks2device((
            (ParallelMode.GLOBAL, ctx.get_global_rank()),
            (ParallelMode.TENSOR, local(ParallelMode.TENSOR)),
            (ParallelMode.PIPELINE, local(ParallelMode.PIPELINE)),
            (ParallelMode.DATA, local(ParallelMode.DATA)),
        ))
        n_gpus = torch.cuda.device_count() if torch.cuda.is_available() else 0
        local_device = (ctx.local_rank % n_gpus) if n_gpus > 0 else ctx.local_rank
        module.parallel_metadata = ParallelMetadata(device=device, local_device=local_device)
        if not getattr(module, "_pg_to_patched", False):
            module._pg_orig_to = module.to
            module.to = partial(_to_device, module)
            module.cuda = partial(_to_cuda, module)
            module._pg_to_patched = True


def _resolve_device(module: nn.Module, device) -> torch.device:
    if isinstance(device, torch.device):
        dev = device
    else:
        dev = torch.device("cuda" if device in ("cuda", "gpu") else device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", module.parallel_metadata.local_device)
    return dev


def _to_device(self: nn.Module, device=None, *args, **kwargs):
    """Move a parallelized module to this rank's device; returns the module (reference returned None)."""
    if device is None or isinstance(device, torch.dtype):
        out = self._pg_orig_to(device, *args, **kwargs) if device is not None else self._pg_orig_to(*args, **kwargs)
    else:
        dev = _resolve_device(self, device)
        if dev.type == "cuda":
            torch.cuda.set_device(dev)
        out = self._pg_orig_to(dev, *args, **kwargs)
    flat = getattr(self, "_flat_state", None)
    if flat is not None:
        flat.rebind()
    for hook in getattr(self, "_pg_after_move_hooks", []):
        hook(self)
    return out


def _to_cuda(self: nn.Module, device: Optional[Union[int, torch.device]] = None):
    if device is None:
        return _to_device(self, "cuda")
    if isinstance(device, int):
        return _to_device(self, torch.device("cuda", device))
    return _to_device(self, device)
