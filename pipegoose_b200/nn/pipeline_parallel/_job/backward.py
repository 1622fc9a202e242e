"""Backward job of the job runtime (parity: reference nn/pipeline_parallel/_job/backward.py:19-160).

The package of a backward job carries ``d loss / d stage_output``; the job differentiates the saved
stage output w.r.t. the saved stage input (parameters accumulate ``.grad`` / ``main_grad`` as usual)
and returns ``d loss / d stage_input`` for the previous partition."""
from __future__ import annotations

import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.pipeline_parallel import queue as Q
from pipegoose_b200.nn.pipeline_parallel._comm import send_package
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback
from pipegoose_b200.nn.pipeline_parallel._job.job import Job
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._package import Package
from pipegoose_b200.nn.pipeline_parallel.exception import PipelineGradientFlowError


class _SaveGradLossFunction(torch.autograd.Function):
    """Identity whose backward parks the incoming gradient in the grad-loss store and stops autograd
    there: the last stage's ``loss.backward()`` then only *records* d loss / d output, and the backward
    jobs replay it through the pipeline."""

    @staticmethod
    def forward(ctx, key, metadata, tensor):
        ctx.pg_key, ctx.pg_metadata = key, metadata
        return tensor.detach().clone()

    @staticmethod
    def backward(ctx, grad_output):
        Q.save_grad_loss(grad_output, ctx.pg_key[0], ctx.pg_key[1], ctx.pg_metadata)
        return None, None, None


def save_grad_loss(package: Package) -> Package:
    """The package with its data behind a recording identity: whatever loss is computed from ``package.data``, its
    ``backward()`` parks d loss / d data in the grad-loss store (``queue.get_grad_loss``) instead of flowing into the
    stage (parity: reference _job/backward.py:17-37 — a Package in, the same Package out)."""
    m = package.metadata
    # a fresh leaf: the stage's own graph must stay untouched (and unfreed) until its backward job runs
    leaf = package.data.detach().requires_grad_(True)
    package.data = _SaveGradLossFunction.apply((m.microbatch_idx, m.partition_idx), m, leaf)
    return package


class BackwardJob(Job):
    def run_compute(self):
        m = self.input.metadata
        grad_output = self.input.data
        out = Q.get_output_activations(m.microbatch_idx, m.partition_idx, is_pipeline=True)
        x = Q.get_input_activations(m.microbatch_idx, m.partition_idx)
        if not (isinstance(out, torch.Tensor) and out.requires_grad):
            raise PipelineGradientFlowError("the saved stage output does not require grad")
        torch.autograd.backward(out, grad_tensors=grad_output.to(out.dtype))
        Q.SavedActivation.get_saved_activations((m.microbatch_idx, m.partition_idx))  # release
        return x.grad if isinstance(x, torch.Tensor) and x.requires_grad else None


class CreateBackwardOutputPackageCallback(Callback):
    order = 0

    def __init__(self, parallel_context, pipeline_context=None):
        self.parallel_context = parallel_context

    def after_compute(self):
        ctx = self.parallel_context
        m = self.job.input.metadata
        is_first = ctx.is_first_rank(ParallelMode.PIPELINE)
        dst = ctx.get_global_rank() if is_first else ctx.get_prev_global_rank(ParallelMode.PIPELINE)
        meta = self.job.input.clone_metadata(job_type=JobType.BACKWARD, partition_idx=m.partition_idx - (0 if is_first else 1),
                                             src=ctx.get_global_rank(), dst=dst)
        self.job.output = Package(self.job.output, meta)


class SendBackwardPackageCallback(Callback):
    order = 5

    def __init__(self, parallel_context):
        self.parallel_context = parallel_context

    def after_compute(self):
        if not self.parallel_context.is_first_rank(ParallelMode.PIPELINE) and self.job.output.data is not None:
            send_package(self.job.output, self.parallel_context)
        self.job.mark_done()
