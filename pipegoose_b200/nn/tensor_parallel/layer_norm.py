"""LayerNorm with replicated parameters (parity: reference nn/tensor_parallel/layer_norm.py:8-25);
runs the fused sm_100a LayerNorm kernel on CUDA bf16 inputs."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext


class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps: float = 1e-5, bias: bool = True, parallel_context: ParallelContext = None):
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = tuple(normalized_shape)
        self.eps = eps
        self.parallel_context = parallel_context
        self.weight = nn.Parameter(torch.ones(self.normalized_shape))
        self.bias = nn.Parameter(torch.zeros(self.normalized_shape)) if bias else None

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if input.is_cuda and input.dtype == torch.bfloat16 and self.bias is not None and len(self.normalized_shape) == 1 \
                and input.shape[-1] % 8 == 0:
            from pipegoose_b200.models.bloom import fused_layer_norm

            return fused_layer_norm(input, self.weight, self.bias, self.eps)
        return F.layer_norm(input, self.normalized_shape, self.weight, self.bias, self.eps)
