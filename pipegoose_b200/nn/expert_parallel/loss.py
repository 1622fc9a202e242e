"""Task loss plus the routers' auxiliary losses (parity: reference nn/expert_parallel/loss.py:8-29).

    loss_fn = ExpertLoss(torch.nn.CrossEntropyLoss(), aux_weight=0.01, z_weight=0.1)
    loss = loss_fn(logits, targets)      # = CE + 0.01 * sum(load-balancing losses) + 0.1 * sum(router z-losses)

The auxiliary terms are whatever the expert layers pushed into :class:`ExpertContext` since the last call; calling the
loss consumes them, so every forward pass is counted exactly once."""
from typing import Callable, List

import torch

from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext


def _weighted_sum(weight: float, terms: List):
    """``weight * sum(terms)``; terms are scalar tensors (what the routers push) or plain numbers (the reference's
    tests push floats)."""
    if not terms:
        return None
    tensors = [t.float().reshape(()) for t in terms if isinstance(t, torch.Tensor)]
    numbers = sum(float(t) for t in terms if not isinstance(t, torch.Tensor))
    total = torch.stack(tensors).sum() + numbers if tensors else torch.tensor(numbers)
    return weight * total


class ExpertLoss:
    def __init__(self, loss_func: Callable, aux_weight: float = 0.01, z_weight: float = 0.1):
        self.loss_func = loss_func
        self.aux_weight = aux_weight
        self.z_weight = z_weight

    # peeking does not consume
    @property
    def aux_loss(self) -> List[torch.Tensor]:
        return ExpertContext.get_instance().aux_loss

    @property
    def z_loss(self) -> List[torch.Tensor]:
        return ExpertContext.get_instance().z_loss

    def __call__(self, *args, **kwargs) -> torch.Tensor:
        total = self.loss_func(*args, **kwargs)
        store = ExpertContext.get_instance()
        for weight, terms in ((self.aux_weight, store.pop_all_aux_loss()), (self.z_weight, store.pop_all_z_loss())):
            extra = _weighted_sum(weight, terms)
            if extra is not None:
                total = total + extra.to(device=total.device, dtype=total.dtype)
        return total
