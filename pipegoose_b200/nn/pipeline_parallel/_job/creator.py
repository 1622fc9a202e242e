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
from pipegoose_b200.nn.pipeline_parallel import queue as Q
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback
from pipegoose_b200.nn.pipeline_parallel._job.job import Job
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._package import Package


class JobCreator(ABC):
    @abstractmethod
    def create(self) -> Job:
        raise NotImplementedError


class ScheduleBackwardJobCallback(Callback):
    """The backward trigger as a forward-job callback (parity: reference _job/creator.py:36-61).  On the last stage
    it swaps the job's output for the recording wrapper of :func:`schedule_backward_execution` and keeps it in the
    scheduled-activations store, so whatever loss the caller computes from the job output, ``loss.backward()`` parks
    d loss / d output in the grad-loss store where the backward jobs pick it up.  The reference runs the ENTIRE
    backward schedule from inside that autograd hook (with sleeps and RPC); here the engine drives the backward tasks
    from its static table, so every micro-batch is treated alike and no hook blocks autograd."""

    order = 3

    def __init__(self, pipeline_context=None, parallel_context=None):
        if parallel_context is None:
            parallel_context = getattr(pipeline_context, "parallel_context", None)
        if parallel_context is None:
            from pipegoose_b200.distributed.parallel_context import ParallelContext

            parallel_context = ParallelContext.get_context()
        self.parallel_context = parallel_context
        self.pipeline_context = pipeline_context

    def after_compute(self):
        from pipegoose_b200.distributed.parallel_mode import ParallelMode

        if not self.parallel_context.is_last_rank(ParallelMode.PIPELINE):
            return
        package = self.job.output
        meta = package.metadata
        recorded = schedule_backward_execution(package).data
        Q._SAVED_SCHEDULED_ACTIVATIONS[(meta.microbatch_idx, meta.partition_idx)] = recorded
        self.job.output = Package(recorded, meta)


class _ForwardJobCreator(JobCreator):
    @classmethod
    def create(cls, function: Callable, package: Package, parallel_context, pipeline_context=None,
               schedule_backward: bool = False) -> ForwardJob:
        cbs = [
            SaveInputActivationsCallback(),
            CreateForwardOutputPackageCallback(parallel_context, pipeline_context),
            SaveBufferForBackwardCallback(),
            SendForwardPackageCallback(parallel_context),
            ConfirmCompleteATaskToProgressTracker(parallel_context),
        ]
        if schedule_backward:
            cbs.append(ScheduleBackwardJobCallback(pipeline_context, parallel_context))
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


def create_job(function: Callable, package: Package, parallel_context, pipeline_context=None,
               schedule_backward: bool = False) -> Job:
    """Forward or backward job for ``package`` with the standard callbacks (``schedule_backward``: forward jobs of the
    last stage also get :class:`ScheduleBackwardJobCallback`)."""
    assert isinstance(package, Package), f"package must be a Package, got {type(package)}"
    creator = _CREATORS[package.metadata.job_type]
    if package.metadata.job_type is JobType.FORWARD:
        return creator.create(function, package, parallel_context, pipeline_context, schedule_backward=schedule_backward)
    return creator.create(function, package, parallel_context, pipeline_context)


def schedule_backward_job(package: Package, pipeline_context=None, parallel_context=None) -> Package:
    """Event-driven backward trigger (parity: reference _job/creator.py:162-180): the package's data passes through an
    identity whose backward builds the :class:`BackwardJob` of this (micro-batch, partition) from the incoming
    gradient and puts it into ``JobQueue.PENDING_JOBS`` for the worker pool.  The data is cut from the stage's graph
    first (the backward job replays the gradient through the saved stage output; letting autograd also run on through
    the stage would differentiate it twice).
    The engines of this library drive backward from static schedule tables and do not need it; it is kept for code
    written against the reference's job runtime."""
    import torch

    if parallel_context is None:
        parallel_context = getattr(pipeline_context, "parallel_context", None)
    if parallel_context is None:
        from pipegoose_b200.distributed.parallel_context import ParallelContext

        parallel_context = ParallelContext.get_context()

    class _Trigger(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.view_as(x)

        @staticmethod
        def backward(ctx, grad):
            job = create_job(None, Package(grad.detach(), package.clone_metadata(job_type=JobType.BACKWARD)),
                             parallel_context, pipeline_context)
            Q.JobQueue.PENDING_JOBS.put(job)
            return grad

    package.data = _Trigger.apply(package.data.detach().requires_grad_(True))
    return package


def schedule_backward_execution(package: Package, pipeline_context=None):
    """Wrap the last stage's output so that ``loss.backward()`` records d loss / d output in the grad-loss
    store (``queue.get_grad_loss``) instead of flowing into the stage: backward jobs start from it.  Returns the
    package (its ``data`` is the wrapped tensor), like the reference."""
    return save_grad_loss(package)
