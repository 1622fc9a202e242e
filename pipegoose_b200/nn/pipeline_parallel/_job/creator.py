"""Build the job for a received package (parity: reference nn/pipeline_parallel/_job/creator.py:28-277).

``create_job(function, package, parallel_context, pipeline_context)`` dispatches on the package's
``job_type`` and attaches the default callback chain.  The reference additionally hides the whole
backward *schedule* inside an autograd hook of the last micro-batch (``schedule_backward_execution``);
here schedules are static tables run by the engine, and :func:`schedule_backward_execution` only records
the loss gradient so that backward jobs can be created from it."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable

from pipegoose_b200.nn.pipeline_parallel._job.backward import (
    BackwardJob,
    CreateBackwardOutputPackageCallback,
    SendBackwardPackageCallback,
    save_grad_loss,
)
from pipegoose_b200.nn.pipeline_parallel._job.forward import (
    ConfirmCompleteATaskToProgressTracker,
    CreateForwardOutputPackageCallback,
    ForwardJob,
    SaveBufferForBackwardCallback,
    SaveInputActivationsCallback,
    SendForwardPackageCallback,
)
from pipegoose_b200.nn.pipeline_parallel._job.job import Job
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._package import Package


class JobCreator(ABC):
    @abstractmethod
    def create(self) -> Job:
        raise NotImplementedError


class _ForwardJobCreator(JobCreator):
    @classmethod
    def create(cls, function: Callable, package: Package, parallel_context, pipeline_context=None) -> ForwardJob:
        cbs = [
            SaveInputActivationsCallback(),
            CreateForwardOutputPackageCallback(parallel_context, pipeline_context),
            SaveBufferForBackwardCallback(),
            SendForwardPackageCallback(parallel_context),
            ConfirmCompleteATaskToProgressTracker(parallel_context),
        ]
        return ForwardJob(function, package, cbs)


class _BackwardJobCreator(JobCreator):
    @classmethod
    def create(cls, function: Callable, package: Package, parallel_context, pipeline_context=None) -> BackwardJob:
        cbs = [
            CreateBackwardOutputPackageCallback(parallel_context, pipeline_context),
            SendBackwardPackageCallback(parallel_context),
            ConfirmCompleteATaskToProgressTracker(parallel_context),
        ]
        return BackwardJob(function, package, cbs)


_CREATORS = {JobType.FORWARD: _ForwardJobCreator, JobType.BACKWARD: _BackwardJobCreator}


def create_job(function: Callable, package: Package, parallel_context, pipeline_context=None) -> Job:
    """Forward or backward job for ``package`` with the standard callbacks."""
    assert isinstance(package, Package), f"package must be a Package, got {type(package)}"
    return _CREATORS[package.metadata.job_type].create(function, package, parallel_context, pipeline_context)


def schedule_backward_execution(package: Package):
    """Wrap the last stage's output so that ``loss.backward()`` records d loss / d output in the grad-loss
    store (``queue.get_grad_loss``) instead of flowing into the stage: backward jobs start from it."""
    return save_grad_loss(package)
