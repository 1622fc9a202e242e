"""``_PipelineEngine``: the engine-style constructor of the reference (nn/pipeline_parallel/pipeline.py:8-46, a stub
there: ``parallelize()`` is ``pass`` and ``forward`` references undefined names).  Here it is a working front door to
the same machinery as :class:`PipelineParallel`: pick a schedule, choose how many worker threads the job runtime may
use, get back the module whose ``forward`` runs the pipeline.

``num_concurrent`` / ``max_concurrent`` are validated and kept for API parity.  Both runtimes execute the jobs of a
stage in clock order (the static runtime on the calling thread from a precomputed schedule table, the job runtime on
one worker thread fed by the progress tracker), because a stage's micro-batches share one CUDA stream and one p2p
channel per neighbour; more worker threads would not add concurrency, only reordering hazards.
"""
from __future__ import annotations

from torch import nn

from pipegoose_b200.constants import PIPELINE_MAX_WORKERS, PIPELINE_MIN_WORKERS
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType, get_scheduler


class _PipelineEngine:
    def __init__(self, module: nn.Module, num_concurrent: int = PIPELINE_MIN_WORKERS,
                 max_concurrent: int = PIPELINE_MAX_WORKERS, scheduler: SchedulerType = SchedulerType.GPIPE,
                 parallel_context: ParallelContext = None, num_microbatches: int = 1, runtime: str = "static"):
        if not isinstance(module, nn.Module):
            raise TypeError(f"module must be an nn.Module, got {type(module).__name__}")
        if not isinstance(parallel_context, ParallelContext):
            raise TypeError("a ParallelContext is required")
        for name, value in (("num_concurrent", num_concurrent), ("max_concurrent", max_concurrent)):
            if not isinstance(value, int) or value < 1:
                raise ValueError(f"{name} must be a positive int, got {value!r}")
        if num_concurrent > max_concurrent:
            raise ValueError("num_concurrent must not exceed max_concurrent")
        self.module = module
        self.num_concurrent = num_concurrent
        self.max_concurrent = max_concurrent
        self.scheduler_type = scheduler
        self.scheduler = get_scheduler(scheduler)   # the scheduler class, like the reference keeps it
        self.parallel_context = parallel_context
        self.num_microbatches = num_microbatches
        self.runtime = runtime
        self._wrapper = None

    def parallelize(self) -> nn.Module:
        from pipegoose_b200.nn.pipeline_parallel.pipeline_parallel import PipelineParallel

        self._wrapper = PipelineParallel(self.module, self.num_microbatches, self.parallel_context,
                                         scheduler_type=self.scheduler_type, runtime=self.runtime)
        return self._wrapper.parallelize()

    def forward(self, *args, **kwargs):
        """Run one pipelined step on the parallelized module (``parallelize()`` first)."""
        if self._wrapper is None:
            self.parallelize()
        return self.module(*args, **kwargs)

    __call__ = forward

    def deparallelize(self) -> nn.Module:
        return self._wrapper.deparallelize() if self._wrapper is not None else self.module
