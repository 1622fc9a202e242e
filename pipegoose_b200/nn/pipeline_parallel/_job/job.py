"""Unit of pipeline work executed by a worker (parity: reference nn/pipeline_parallel/_job/job.py:9-125).

Differences from the reference: a failing job is marked ``FAILED`` (the reference declares the
status but never sets it, so a raising job hung the pipeline) and keeps the exception in
``job.error``; keys come from a process-wide counter plus a random suffix instead of 15 random
characters, so logs sort by creation order."""
from __future__ import annotations

import itertools
import random
import threading
import string
from abc import ABC, abstractmethod
from enum import Enum, auto
from typing import Callable, List, NewType, Optional

from pipegoose_b200.constants import JOB_KEY_LENGTH
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback, CallbackEvent
from pipegoose_b200.nn.pipeline_parallel._package import Package


class JobStatus(Enum):
    PENDING = auto()    # created, waiting in a queue
    EXECUTING = auto()
    EXECUTED = auto()   # computed, output not yet handed to the next stage
    DONE = auto()       # computed and output delivered
    FAILED = auto()


PartitionKey = NewType("PartitionKey", str)
_COUNTER = itertools.count()


def _make_key() -> str:
    head = f"{next(_COUNTER):06d}"
    tail = "".join(random.choice(string.ascii_letters + string.digits) for _ in range(JOB_KEY_LENGTH - len(head)))
    return head + tail


class Job(ABC):
    def __init__(self, function: Callable, input: Package, cbs: List[Callback] = ()):
        self.function = function
        self.input = input
        self.cbs: List[Callback] = []
        self._status = JobStatus.PENDING
        self._output: Optional[Package] = None
        self._key = _make_key()
        self.error: Optional[BaseException] = None
        self._finished = threading.Event()  # set when compute() returned or raised
        self.add_cbs(cbs)
        self._run_callback(CallbackEvent.AFTER_CREATE)

    # ------------------------------------------------------------------ state
    @property
    def status(self) -> JobStatus:
        return self._status

    @property
    def key(self) -> str:
        return self._key

    @property
    def output(self) -> Optional[Package]:
        return self._output

    @output.setter
    def output(self, value: Optional[Package]):
        self._output = value

    def mark_done(self):
        self._status = JobStatus.DONE

    # ------------------------------------------------------------------ execution
    def compute(self) -> Optional[Package]:
        try:
            self._status = JobStatus.EXECUTING
            self._run_callback(CallbackEvent.BEFORE_COMPUTE)
            self._output = self.run_compute()
            self._status = JobStatus.EXECUTED
            self._run_callback(CallbackEvent.AFTER_COMPUTE)
            return self._output
        except BaseException as e:
            self._status = JobStatus.FAILED
            self.error = e
            self._run_callback(CallbackEvent.ON_FAILURE)
            raise
        finally:
            self._finished.set()

    def wait(self, timeout: Optional[float] = None) -> bool:
        """Block until a worker finished (or failed) this job."""
        return self._finished.wait(timeout)

    @abstractmethod
    def run_compute(self):
        """The actual computation of this job."""

    # ------------------------------------------------------------------ callbacks
    def add_cbs(self, cbs):
        for cb in cbs:
            self.add_cb(cb)

    def remove_cbs(self, cbs):
        for cb in list(cbs):
            self.remove_cb(cb)

    def add_cb(self, cb):
        if isinstance(cb, type):
            cb = cb()
        assert isinstance(cb, Callback), f"cb must be a Callback, got {type(cb)}"
        cb.job = self
        self.cbs.append(cb)

    def remove_cb(self, cb):
        if isinstance(cb, type):
            self.cbs = [x for x in self.cbs if not isinstance(x, cb)]
        elif cb in self.cbs:
            self.cbs.remove(cb)

    def _run_callback(self, event: CallbackEvent):
        assert isinstance(event, CallbackEvent), f"expected a CallbackEvent, got {type(event)}"
        for cb in sorted(self.cbs, key=lambda c: c.order):
            method = getattr(cb, event.value, None)
            if method is not None:
                method()
