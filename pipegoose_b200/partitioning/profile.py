"""Per-child memory profile of a module (parity: reference partitioning/profile.py:9-49), the input
of balance-by-memory pipeline partitioning.  Works on CUDA (allocated-bytes delta) and on CPU
(parameter + output bytes)."""
from __future__ import annotations

from typing import List

import torch
from torch import nn


def _tensor_bytes(obj) -> int:
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (list, tuple)):
        return sum(_tensor_bytes(o) for o in obj)
    if isinstance(obj, dict):
        return sum(_tensor_bytes(o) for o in obj.values())
    return 0


class ProfileStrategy:
    """What a partitioning cost model looks like: ``profile(input) -> [cost per child]`` (parity: reference
    partitioning/profile.py:9-16)."""

    def __init__(self, module: nn.Module, device=None):
        self.module = module
        self.device = device

    def profile(self, input: torch.Tensor) -> List[int]:
        raise NotImplementedError


class ProfileByMemory(ProfileStrategy):
    def __init__(self, module: nn.Sequential, device=None):
        self.module = module
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))

    @torch.no_grad()
    def profile(self, input: torch.Tensor) -> List[int]:
        sizes = []
        x = input.to(self.device)
        for layer in self.module:
            layer = layer.to(self.device)
            param_bytes = sum(p.numel() * p.element_size() for p in layer.parameters())
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
                before = torch.cuda.memory_allocated(self.device)
                x = layer(x)
                torch.cuda.synchronize(self.device)
                act_bytes = max(0, torch.cuda.memory_allocated(self.device) - before)
            else:
                x = layer(x)
                act_bytes = _tensor_bytes(x)
            sizes.append(int(param_bytes + act_bytes))
        return sizes
