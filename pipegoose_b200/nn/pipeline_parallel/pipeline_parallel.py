"""``PipelineParallel`` wrapper (parity: reference nn/pipeline_parallel/pipeline_parallel.py:13-50):
keep this rank's partition of the module and route ``module.forward`` through the pipeline engine."""
from __future__ import annotations

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.nn.parallel import Parallel
from pipegoose_b200.nn.pipeline_parallel._utils import get_partition_idx
from pipegoose_b200.nn.pipeline_parallel.partitioner import UniformPartitioner
from pipegoose_b200.nn.pipeline_parallel.pipeline_context import PipelineContext
from pipegoose_b200.nn.pipeline_parallel.pipeline_engine import PipelineEngine
from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType, get_scheduler


class PipelineParallel(Parallel):
    def __init__(self, module: nn.Module, num_microbatches: int, parallel_context: ParallelContext,
                 scheduler_type: SchedulerType = SchedulerType.ONE_F_ONE_B):
        super().__init__(module, parallel_context)
        self.num_microbatches = num_microbatches
        self.scheduler_type = scheduler_type

    @torch.no_grad()
    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        if ctx.pipeline_parallel_size > 1:
            partitions = UniformPartitioner(module, ctx).split(["input_ids"])
            stage = partitions[get_partition_idx(ctx)]
            scheduler = get_scheduler(self.scheduler_type)(self.num_microbatches, ctx.pipeline_parallel_size)
            pipeline_context = PipelineContext(scheduler, ctx)
            engine = PipelineEngine(stage, scheduler, ctx, pipeline_context, full_module=module)
            _drop_foreign_parameters(module, stage)
            module._pg_pipeline_stage = stage
            module._pg_pipeline_engine = engine
            module.forward = engine.run
            self._save_metadata(module, ctx)
        return module

    def deparallelize(self) -> nn.Module:
        raise NotImplementedError


def _drop_foreign_parameters(module: nn.Module, stage: nn.Module):
    """Free the parameters of the other stages (they stay registered but become empty)."""
    keep = {id(p) for p in stage.parameters()}
    for p in module.parameters():
        if id(p) not in keep:
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
            p.requires_grad_(False)
