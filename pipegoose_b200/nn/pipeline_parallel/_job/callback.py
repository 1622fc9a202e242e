"""Job life-cycle callbacks (parity: reference nn/pipeline_parallel/_job/callback.py:5-30).

A :class:`CallbackEvent`'s value is the name of the hook it triggers, so dispatch is ``callback.handle(event)``; a
job runs the callbacks of an event in ascending ``order``.  ``ON_FAILURE`` does not exist in the reference (a job that
raises there leaves its status at EXECUTING and the pipeline hangs).
"""
from __future__ import annotations

from enum import Enum


class CallbackEvent(str, Enum):
    AFTER_CREATE = "after_create"
    BEFORE_COMPUTE = "before_compute"
    AFTER_COMPUTE = "after_compute"
    ON_FAILURE = "on_failure"

    @property
    def hook(self) -> str:
        return self.value


class Callback:
    """Attach with ``job.add_cb(MyCallback)``; ``self.job`` is the job it belongs to."""

    order: int = 0
    job = None

    def handle(self, event: CallbackEvent):
        return getattr(self, CallbackEvent(event).hook)()

    @property
    def name(self) -> str:
        return type(self).__name__

    # hooks: override what is needed ------------------------------------------------------------------
    def after_create(self): ...

    def before_compute(self): ...

    def after_compute(self): ...

    def on_failure(self): ...
