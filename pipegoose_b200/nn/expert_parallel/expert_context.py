"""Process-wide collector of the routers' auxiliary losses (parity: reference
nn/expert_parallel/expert_context.py:7-32)."""
from __future__ import annotations

from typing import List, Optional

import torch


class ExpertContext:
    _instance: Optional["ExpertContext"] = None

    def __init__(self):
        self.aux_loss: List[torch.Tensor] = []
        self.z_loss: List[torch.Tensor] = []

    def push_aux_loss(self, aux_loss: torch.Tensor):
        self.aux_loss.append(aux_loss)

    def pop_all_aux_loss(self) -> List[torch.Tensor]:
        out, self.aux_loss = self.aux_loss, []
        return out

    def push_z_loss(self, z_loss: torch.Tensor):
        self.z_loss.append(z_loss)

    def pop_all_z_loss(self) -> List[torch.Tensor]:
        out, self.z_loss = self.z_loss, []
        return out

    @classmethod
    def get_instance(cls) -> "ExpertContext":
        if cls._instance is None:
            cls._instance = ExpertContext()
        return cls._instance
