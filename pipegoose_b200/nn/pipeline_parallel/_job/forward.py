"""Forward job of the job runtime and its callbacks (parity: reference nn/pipeline_parallel/_job/forward.py:14-133)."""
from __future__ import annotations

import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.pipeline_parallel import queue as Q
from pipegoose_b200.nn.pipeline_parallel._comm import send_package
from pipegoose_b200.nn.pipeline_parallel._job.callback import Callback
from pipegoose_b200.nn.pipeline_parallel._job.job import Job
from pipegoose_b200.nn.pipeline_parallel._package import Package


class ForwardJob(Job):
    """Run the stage function on the package's data under the package's grad mode."""

    def run_compute(self):
        data = self.input.data
        training = self.input.metadata.training
        with torch.set_grad_enabled(training.is_training and training.is_grad_enabled):
            if isinstance(data, torch.Tensor):
                return self.function(data)
            if isinstance(data, (tuple, list)):
                return self.function(*data)
            if isinstance(data, dict):
                return self.function(**data)
            raise TypeError(f"unsupported package payload: {type(data)}")


class SaveInputActivationsCallback(Callback):
    """Keep the stage input: the backward job returns its gradient to the previous stage."""

    order = -1

    def before_compute(self):
        m = self.job.input.metadata
        x = self.job.input.data
        if isinstance(x, torch.Tensor) and x.is_floating_point() and m.training.is_training:
            x = x.detach().requires_grad_(True)
            self.job.input.data = x
        Q.save_input_activations(x, m.microbatch_idx, m.partition_idx)


class CreateForwardOutputPackageCallback(Callback):
    """Save the output for backward and wrap it into the package for the next partition."""

    order = 0

    def __init__(self, parallel_context, pipeline_context=None):
        self.parallel_context = parallel_context
        self.pipeline_context = pipeline_context

    def after_compute(self):
        ctx = self.parallel_context
        m = self.job.input.metadata
        out = self.job.output
        Q.save_output_activations(out, m.microbatch_idx, m.partition_idx)
        is_last = ctx.is_last_rank(ParallelMode.PIPELINE)
        if is_last:
            # the last partition's output stays where it is (it seeds this partition's own backward package): routing
            # fields unchanged, like the reference (tests/nn/pipeline_parallel/job/test_forward.py:60-73)
            meta = self.job.input.clone_metadata()
        else:
            meta = self.job.input.clone_metadata(partition_idx=m.partition_idx + 1, src=ctx.get_global_rank(),
                                                 dst=ctx.get_next_global_rank(ParallelMode.PIPELINE))
        payload = out.detach() if isinstance(out, torch.Tensor) and not is_last else out
        self.job.output = Package(payload, meta)


class SaveBufferForBackwardCallback(Callback):
    """Kept for API parity (the reference's version is a no-op with a misspelt hook name)."""

    order = 1


class SendForwardPackageCallback(Callback):
    order = 5

    def __init__(self, parallel_context):
        self.parallel_context = parallel_context

    def after_compute(self):
        if not self.parallel_context.is_last_rank(ParallelMode.PIPELINE):
            send_package(self.job.output, self.parallel_context)
        self.job.mark_done()


class ConfirmCompleteATaskToProgressTracker(Callback):
    """Tell the progress tracker (if one is installed) that (microbatch, partition) finished."""

    order = 6

    def __init__(self, parallel_context):
        self.parallel_context = parallel_context

    def after_compute(self):
        from pipegoose_b200.nn.pipeline_parallel.sync.handshake import get_progress_tracker

        tracker = get_progress_tracker()
        if tracker is not None and tracker.is_initiated():
            m = self.job.input.metadata
            tracker.confirm((m.microbatch_idx, m.partition_idx))
