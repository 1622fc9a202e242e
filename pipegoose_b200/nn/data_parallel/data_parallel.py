"""Data parallelism (parity: reference nn/data_parallel/data_parallel.py:13-43).

``DataParallel(module, parallel_context).parallelize()`` returns the same module.  Where the
reference registers one blocking all-reduce per parameter, this wrapper:

* broadcasts the parameters from data-parallel rank 0 once, so replicas start identical even
  when seeds differ (the reference relies on identical seeds);
* installs a :class:`GradReducer`: gradients live in one flat fp32 buffer, are averaged bucket
  by bucket while backward is still running, and expert parameters (``param.is_expert``) use
  the EXPERT_DATA group;
* supports ``module.no_sync()`` for gradient accumulation.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from pipegoose_b200.constants import BUCKET_SIZE_MB
from pipegoose_b200.core.grad_reducer import GradReducer
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.parallel import Parallel


class DataParallel(Parallel):
    def __init__(self, module: nn.Module, parallel_context: ParallelContext, bucket_size_mb: float = BUCKET_SIZE_MB,
                 broadcast_parameters: bool = True):
        super().__init__(module, parallel_context)
        self.bucket_size_mb = bucket_size_mb
        self.broadcast_parameters = broadcast_parameters

    @torch.no_grad()
    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        if ctx.data_parallel_size > 1:
            reducer = GradReducer(module, ctx, self.bucket_size_mb)
            module._pg_grad_reducer = reducer
            module.no_sync = reducer.no_sync
            module._pg_dp_build_hook = module.register_forward_pre_hook(_build_on_first_forward)
            reducer.ensure_built = lambda: _build_on_first_forward(module, None)
            for p in module.parameters():   # lets an optimizer created before the first forward find (and build) it
                p._pg_dp_reducer = reducer
            if self.broadcast_parameters:
                module._pg_needs_param_broadcast = True
            self._save_metadata(module, ctx)
        return module

    def deparallelize(self) -> nn.Module:
        """Detach the gradient reducer: hooks removed, ``no_sync`` and the gradient-ready notifications handed back to
        the tensor-parallel partial-gradient sync when the module has one (TP without DP still has to sum the
        gradients of the TP-replicated parameters)."""
        module = self.module
        for p in module.parameters():
            h = getattr(p, "_pg_autograd_hook", None)
            if h is not None:
                h.remove()
                del p._pg_autograd_hook
            if hasattr(p, "_pg_grad_ready"):
                del p._pg_grad_ready
            if hasattr(p, "_pg_dp_reducer"):
                del p._pg_dp_reducer
        h = getattr(module, "_pg_dp_build_hook", None)
        if h is not None:
            h.remove()
            del module._pg_dp_build_hook
        if hasattr(module, "_pg_grad_reducer"):
            del module._pg_grad_reducer
        tp_sync = getattr(module, "_pg_tp_grad_sync", None)
        if tp_sync is not None:
            for p in tp_sync.params:
                p._pg_grad_ready = tp_sync._on_ready
            module.no_sync = tp_sync.no_sync
        elif "no_sync" in module.__dict__:
            del module.no_sync
        return module


def _build_on_first_forward(module: nn.Module, _inputs):
    reducer: GradReducer = module._pg_grad_reducer
    if reducer.flat is None:
        reducer.build()
        if getattr(module, "_pg_needs_param_broadcast", False):
            _broadcast_parameters(module, reducer)
            module._pg_needs_param_broadcast = False


def _broadcast_parameters(module: nn.Module, reducer: GradReducer):
    ctx = reducer.ctx
    group = ctx.get_group(ParallelMode.DATA)
    src = ctx.get_ranks_in_group(ParallelMode.DATA)[0]
    dist.broadcast(reducer.flat.flat_param, src=src, group=group)
    for buf in module.buffers():
        if buf.is_floating_point() or buf.dtype in (torch.int64, torch.int32):
            dist.broadcast(buf, src=src, group=group)
