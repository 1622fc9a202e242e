"""Job life-cycle callbacks (parity: reference nn/pipeline_parallel/_job/callback.py:5-30)."""
from __future__ import annotations

from enum import Enum


class CallbackEvent(Enum):
    AFTER_CREATE = "after_create"
    BEFORE_COMPUTE = "before_compute"
    AFTER_COMPUTE = "after_compute"
    ON_FAILURE = "on_failure"


class Callback:
    """Hook object attached to a :class:`Job`; ``self.job`` is set when it is added.  Callbacks of one
    event run in ascending ``order``."""

    order = 0
    job = None

    @property
    def name(self) -> str:
        return type(self).__name__

    def after_create(self):
        pass

    def before_compute(self):
        pass

    def after_compute(self):
        pass

    def on_failure(self):
        pass
