from abc import ABC, abstractmethod


class BaseDistributedOptimizer(ABC):
    """Contract of a distributed optimizer (parity: reference optim/base_optim.py:4-33)."""

    @property
    @abstractmethod
    def defaults(self):
        raise NotImplementedError

    @property
    @abstractmethod
    def param_groups(self):
        raise NotImplementedError

    @abstractmethod
    def add_param_group(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def load_state_dict(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def state_dict(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def step(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def zero_grad(self, *args, **kwargs):
        raise NotImplementedError
