"""Differentiable communication ops for the replicated-activation tensor-parallel layers
(parity: reference nn/tensor_parallel/_functional.py:15-95).

    broadcast_to_tensor_group : identity forward,  all-reduce backward   (column-parallel input)
    gather_to_tensor_group    : all-gather forward, local slice backward (column-parallel output)
    scatter_to_tensor_group   : local slice forward, all-gather backward (row-parallel input)
    reduce_to_tensor_group    : all-reduce forward, identity backward    (row-parallel output)
"""
from __future__ import annotations

import torch

from pipegoose_b200.distributed.functional import all_gather, all_reduce, scatter
from pipegoose_b200.distributed.parallel_mode import ParallelMode

_TP = ParallelMode.TENSOR


class _Broadcast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, parallel_context):
        ctx.parallel_context = parallel_context
        return tensor

    @staticmethod
    def backward(ctx, grad):
        return all_reduce(grad.contiguous(), parallel_context=ctx.parallel_context, parallel_mode=_TP), None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, dim, parallel_context):
        ctx.dim, ctx.parallel_context = dim, parallel_context
        return all_gather(tensor, dim=dim, parallel_context=parallel_context, parallel_mode=_TP)

    @staticmethod
    def backward(ctx, grad):
        return scatter(grad, dim=ctx.dim, parallel_context=ctx.parallel_context, parallel_mode=_TP).contiguous(), None, None


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, dim, parallel_context):
        ctx.dim, ctx.parallel_context = dim, parallel_context
        return scatter(tensor, dim=dim, parallel_context=parallel_context, parallel_mode=_TP).contiguous()

    @staticmethod
    def backward(ctx, grad):
        return all_gather(grad.contiguous(), dim=ctx.dim, parallel_context=ctx.parallel_context, parallel_mode=_TP), None, None


class _Reduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, parallel_context):
        return all_reduce(tensor.contiguous().clone(), parallel_context=parallel_context, parallel_mode=_TP)

    @staticmethod
    def backward(ctx, grad):
        return grad, None


# keyword names as in the reference (``input=...``), so call sites written against it keep working
def broadcast_to_tensor_group(input, parallel_context):  # noqa: A002
    return _Broadcast.apply(input, parallel_context)


def gather_to_tensor_group(input, dim, parallel_context):  # noqa: A002
    return _Gather.apply(input, dim, parallel_context)


def scatter_to_tensor_group(input, dim, parallel_context):  # noqa: A002
    return _Scatter.apply(input, dim, parallel_context)


def reduce_to_tensor_group(input, parallel_context):  # noqa: A002
    return _Reduce.apply(input, parallel_context)
