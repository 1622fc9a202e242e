"""Top-k routers for Switch-style mixture of experts (parity: reference
nn/expert_parallel/routers.py:12-189).

The gate runs in fp32.  ``RouterOutput.dispatching_order`` is the ``[tokens, E]`` 0/1 routing mask
(after the capacity limit), ``weight`` the gate probabilities of the selected experts
(``prob * mask``), ``aux_loss`` the Switch load-balancing loss
``alpha * E * <mean(mask), mean(prob)>`` and ``z_loss = mean(logsumexp(logits)^2)``.
Differences from the reference (documented quirks Q5): the exploration noise *multiplies* the
logits as in the Switch paper, and the expert layer really applies ``weight`` to the expert
outputs.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum, auto
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


class RouterExplorationNoisePolicy(ABC):
    @abstractmethod
    def sample_like(self, input: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class SwitchNoisePolicy(RouterExplorationNoisePolicy):
    """Multiplicative jitter uniformly distributed in ``[1 - eps, 1 + eps)`` (Switch Transformer, app. A)."""

    def __init__(self, eps: float = 0.1):
        assert eps > 0.0
        self.eps = eps

    def sample_like(self, input: torch.Tensor) -> torch.Tensor:
        noise = torch.rand_like(input)
        return noise * (2 * self.eps) + (1.0 - self.eps)


@dataclass
class RouterOutput:
    dispatching_order: torch.Tensor
    weight: torch.Tensor
    aux_loss: torch.Tensor
    z_loss: torch.Tensor


class RouterType(Enum):
    TOP_1 = auto()
    TOP_2 = auto()


class Router(ABC, nn.Module):
    pass


class _TopKRouter(Router):
    def __init__(self, noise_policy: Optional[RouterExplorationNoisePolicy], top_k: int, num_experts: int, d_model: int,
                 expert_capacity: Optional[Tuple[float, float]] = None, alpha: float = 0.01, eps: float = 0.1):
        super().__init__()
        self.noise_policy = noise_policy
        self.top_k = top_k
        self.num_experts = num_experts
        self.expert_capacity = expert_capacity
        self.alpha = alpha
        self.eps = eps
        self.gate = nn.Linear(d_model, num_experts)

    def _aux_loss(self, router_prob: torch.Tensor, expert_mask: torch.Tensor) -> torch.Tensor:
        tokens_per_expert = expert_mask.float().mean(dim=0)
        prob_per_expert = router_prob.mean(dim=0)
        return self.alpha * self.num_experts * torch.dot(tokens_per_expert, prob_per_expert)

    def _z_loss(self, router_logits: torch.Tensor) -> torch.Tensor:
        return torch.logsumexp(router_logits, dim=-1).square().mean()

    def _expert_capacity(self, total_tokens: int) -> int:
        if self.expert_capacity is None:
            return total_tokens
        factor = self.expert_capacity[0] if self.training else self.expert_capacity[1]
        return max(1, math.ceil(total_tokens / self.num_experts * factor))

    def forward(self, inputs: torch.Tensor) -> RouterOutput:
        x = inputs.reshape(-1, inputs.shape[-1]).to(torch.float32)
        logits = F.linear(x, self.gate.weight.float(), self.gate.bias.float())
        if self.training and self.noise_policy is not None:
            logits = logits * self.noise_policy.sample_like(logits)
        prob = F.softmax(logits, dim=-1)
        _, top_idx = torch.topk(prob, k=self.top_k, dim=-1)
        mask = torch.zeros_like(prob).scatter_(1, top_idx, 1.0)
        aux_loss = self._aux_loss(prob, mask)
        z_loss = self._z_loss(logits)
        if self.expert_capacity is not None:
            capacity = self._expert_capacity(x.shape[0])
            position = torch.cumsum(mask, dim=0) * mask  # 1-based arrival order inside each expert
            mask = mask * (position <= capacity).to(mask.dtype)
        weight = prob * mask
        return RouterOutput(dispatching_order=mask, weight=weight, aux_loss=aux_loss, z_loss=z_loss)


class Top1Router(_TopKRouter):
    def __init__(self, noise_policy, num_experts: int, d_model: int, expert_capacity=None, alpha: float = 0.01,
                 eps: float = 0.1):
        super().__init__(noise_policy, 1, num_experts, d_model, expert_capacity, alpha, eps)


class Top2Router(_TopKRouter):
    def __init__(self, noise_policy, num_experts: int, d_model: int, expert_capacity=None, alpha: float = 0.01,
                 eps: float = 0.1):
        super().__init__(noise_policy, 2, num_experts, d_model, expert_capacity, alpha, eps)
