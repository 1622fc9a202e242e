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
            engine.tied_group, engine.tied_param = _tied_embedding_group(module, ctx)
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


def _tied_embedding_group(module: nn.Module, ctx: ParallelContext):
    """Input embedding (first stage) and lm_head (last stage) share one table: their gradient
    contributions are summed over a 2-rank group after every backward.  Collective (``new_group``):
    every rank creates the group of every pipeline group."""
    import torch.distributed as dist

    from pipegoose_b200.distributed.parallel_mode import ParallelMode

    get_in = getattr(module, "get_input_embeddings", None)
    get_out = getattr(module, "get_output_embeddings", None)
    tied = None
    if get_in is not None and get_out is not None:
        emb, head = get_in(), get_out()
        if emb is not None and head is not None and head.weight is emb.weight:
            tied = emb.weight
    flag = torch.tensor([1 if tied is not None else 0])
    my_group = None
    if int(flag.item()):
        for ranks in ctx.topology.groups(ParallelMode.PIPELINE):
            pair = [ranks[0], ranks[-1]]
            g = dist.new_group(ranks=pair)
            if ctx.get_global_rank() in pair:
                my_group = g
    return my_group, tied
