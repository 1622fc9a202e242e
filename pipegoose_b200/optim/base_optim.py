"""The contract of a distributed optimizer (parity: reference optim/base_optim.py:4-33 — seven abstract stubs).

Instead of abstract stubs the contract is a table checked once, when a subclass is defined: a wrapper that forgets one
of the members fails at import time, not at the first ``step()`` of a long job.
"""
from __future__ import annotations

from typing import Tuple


class BaseDistributedOptimizer:
    #: what user code, LR schedulers and the checkpoint functions touch on an optimizer
    CONTRACT: Tuple[str, ...] = ("defaults", "param_groups", "add_param_group", "load_state_dict", "state_dict", "step",
                                 "zero_grad")

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        absent = [name for name in BaseDistributedOptimizer.CONTRACT if not hasattr(cls, name)]
        if absent:
            raise TypeError(f"{cls.__name__} does not implement the distributed-optimizer contract: missing {absent}")

    def __new__(cls, *args, **kwargs):
        if cls is BaseDistributedOptimizer:
            raise TypeError("BaseDistributedOptimizer is a contract, not an optimizer: subclass it")
        return super().__new__(cls)
