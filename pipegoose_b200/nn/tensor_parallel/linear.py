"""1-D (Megatron-style) tensor-parallel linear layers with the reference's replicated-activation
contract (parity: reference nn/tensor_parallel/linear.py:17-82).

The local GEMMs run on the tcgen05 kernel (``ops.functional.linear``: bias in the epilogue).
These classes keep the reference semantics (``gather_output`` all-gathers the column slices, the
row-parallel output is all-reduced) so that any 🤗 model can be sharded; the fast
sequence-parallel path with fused all-gather->GEMM / GEMM->reduce-scatter kernels is used by
``pipegoose_b200.models`` (see ops/functional.py).
"""
from __future__ import annotations

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.tensor_parallel._functional import (
    broadcast_to_tensor_group,
    gather_to_tensor_group,
    reduce_to_tensor_group,
    scatter_to_tensor_group,
)
from pipegoose_b200.ops.functional import linear as fused_linear


def _tp_size(parallel_context) -> int:
    return parallel_context.get_world_size(ParallelMode.TENSOR) if parallel_context is not None else 1


class ColumnParallelLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, gather_output: bool = False,
                 parallel_context: ParallelContext = None):
        super().__init__()
        world = _tp_size(parallel_context)
        assert out_features % world == 0, "out_features must be divisible by the tensor parallel size"
        self.in_features = in_features
        self.out_features = out_features
        self.gather_output = gather_output
        self.parallel_context = parallel_context
        self.weight = nn.Parameter(torch.empty(out_features // world, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features // world)) if bias else None
        nn.init.normal_(self.weight, std=0.02)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        x = broadcast_to_tensor_group(input, self.parallel_context)
        out = fused_linear(x, self.weight, self.bias)
        if self.gather_output:
            out = gather_to_tensor_group(out, dim=-1, parallel_context=self.parallel_context)
        return out


class RowParallelLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, parallel_context: ParallelContext = None):
        super().__init__()
        world = _tp_size(parallel_context)
        assert in_features % world == 0, "in_features must be divisible by the tensor parallel size"
        self.in_features = in_features
        self.out_features = out_features
        self.parallel_context = parallel_context
        self.weight = nn.Parameter(torch.empty(out_features, in_features // world))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        nn.init.normal_(self.weight, std=0.02)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        x = scatter_to_tensor_group(input, dim=-1, parallel_context=self.parallel_context)
        partial = fused_linear(x, self.weight, None)
        out = reduce_to_tensor_group(partial, self.parallel_context)
        if self.bias is not None:
            out = out + self.bias
        return out
