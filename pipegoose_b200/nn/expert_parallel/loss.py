"""Task loss + router auxiliary losses (parity: reference nn/expert_parallel/loss.py:8-29)."""
from typing import Callable

import torch

from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext


class ExpertLoss:
    def __init__(self, loss_func: Callable, aux_weight: float = 0.01, z_weight: float = 0.1):
        self.loss_func = loss_func
        self.aux_weight = aux_weight
        self.z_weight = z_weight

    @property
    def aux_loss(self):
        return ExpertContext.get_instance().aux_loss

    @property
    def z_loss(self):
        return ExpertContext.get_instance().z_loss

    def __call__(self, *args, **kwargs) -> torch.Tensor:
        loss = self.loss_func(*args, **kwargs)
        ctx = ExpertContext.get_instance()
        aux, z = ctx.pop_all_aux_loss(), ctx.pop_all_z_loss()
        if aux:
            loss = loss + self.aux_weight * sum(aux)
        if z:
            loss = loss + self.z_weight * sum(z)
        return loss
