"""Process-wide collector of the routers' auxiliary losses (parity: reference nn/expert_parallel/expert_context.py:7-32).

Every ``ExpertLayer`` / ``FusedExpertLayer`` forward pushes its load-balancing (aux) and router-z losses here;
``ExpertLoss`` pops them when the task loss is computed.  One store per process, guarded by a lock because pipeline
worker threads may run expert layers concurrently."""
from __future__ import annotations

import threading
from typing import Dict, List, Optional

import torch

_KINDS = ("aux", "z")


class ExpertContext:
    _instance: Optional["ExpertContext"] = None
    _instance_lock = threading.Lock()

    def __init__(self):
        self._lock = threading.Lock()
        self._losses: Dict[str, List[torch.Tensor]] = {k: [] for k in _KINDS}

    # ------------------------------------------------------------------ singleton
    @classmethod
    def get_instance(cls) -> "ExpertContext":
        with cls._instance_lock:
            if cls._instance is None:
                cls._instance = cls()
            return cls._instance

    # ------------------------------------------------------------------ generic store
    def _push(self, kind: str, value: torch.Tensor):
        with self._lock:
            self._losses[kind].append(value)

    def _pop_all(self, kind: str) -> List[torch.Tensor]:
        with self._lock:
            out, self._losses[kind] = self._losses[kind], []
        return out

    def clear(self):
        for kind in _KINDS:
            self._pop_all(kind)

    # ------------------------------------------------------------------ reference API
    @property
    def aux_loss(self) -> List[torch.Tensor]:
        return self._losses["aux"]

    @property
    def z_loss(self) -> List[torch.Tensor]:
        return self._losses["z"]

    def push_aux_loss(self, aux_loss: torch.Tensor):
        self._push("aux", aux_loss)

    def pop_all_aux_loss(self) -> List[torch.Tensor]:
        return self._pop_all("aux")

    def push_z_loss(self, z_loss: torch.Tensor):
        self._push("z", z_loss)

    def pop_all_z_loss(self) -> List[torch.Tensor]:
        return self._pop_all("z")
