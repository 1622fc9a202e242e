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


def _generate_not_pipelined(*args, **kwargs):
    raise NotImplementedError(
        "generate() of this model class is not pipeline-aware.  pipegoose_b200.models (BloomForCausalLM, GPT2LMHeadModel; a "
        "🤗 Bloom converted by TensorParallel(model, ctx, sequence_parallel=True)) generate on pipelined layouts — one "
        "forward-only schedule per token; otherwise call PipelineParallel(...).deparallelize() first.")


class PipelineParallel(Parallel):
    def __init__(self, module: nn.Module, num_microbatches: int, parallel_context: ParallelContext,
                 scheduler_type: SchedulerType = SchedulerType.ONE_F_ONE_B, runtime: str = "static",
                 aux_loss_weight: float = 0.01, z_loss_weight: float = 0.001, partitioner=None):
        """``runtime``: ``"static"`` (default) — schedule tables + batched p2p (pipeline_engine.py); ``"jobs"`` — the
        reference's execution model: jobs from packages, worker threads, progress tracker (job_engine.py, GPipe).
        ``partitioner``: ``None`` (the ``UniformPartitioner``: known families by structure, anything else through its
        ``torch.fx`` graph) or a callable ``(module, parallel_context) -> partitioner with .split()`` — e.g.
        ``lambda m, c: GraphPartitioner(m, c, leaf_modules=(MyBlock,), concrete_args={...})``."""
        super().__init__(module, parallel_context)
        self.partitioner = partitioner
        assert runtime in ("static", "jobs")
        self.num_microbatches = num_microbatches
        self.scheduler_type = scheduler_type if runtime == "static" else SchedulerType.GPIPE
        self.runtime = runtime
        # Switch-MoE stages: weights of the load-balancing / router-z losses each stage back-propagates (static runtime)
        self.aux_loss_weight, self.z_loss_weight = aux_loss_weight, z_loss_weight

    @torch.no_grad()
    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        if ctx.pipeline_parallel_size > 1:
            make = self.partitioner if self.partitioner is not None else UniformPartitioner
            partitions = make(module, ctx).split(["input_ids"])
            assert len(partitions) == ctx.pipeline_parallel_size, "the partitioner must return one stage per pipeline rank"
            stage = partitions[get_partition_idx(ctx)]
            scheduler = get_scheduler(self.scheduler_type)(self.num_microbatches, ctx.pipeline_parallel_size)
            pipeline_context = PipelineContext(scheduler, ctx)
            if self.runtime == "jobs":
                from pipegoose_b200.nn.pipeline_parallel.job_engine import JobPipelineEngine

                engine = JobPipelineEngine(stage, scheduler, ctx, pipeline_context, full_module=module)
            else:
                engine = PipelineEngine(stage, scheduler, parallel_context=ctx, pipeline_context=pipeline_context,
                                        full_module=module)
            engine.aux_loss_weight, engine.z_loss_weight = self.aux_loss_weight, self.z_loss_weight
            engine.tied_group, engine.tied_param = _tied_embedding_group(module, ctx)
            if engine.tied_group is not None and engine.tied_param is not None:
                engine.tied_param._pg_pp_shared = 2  # lives on the first and the last stage (global norms: half each)
            _drop_foreign_parameters(module, stage)
            module._pg_pipeline_stage = stage
            module._pg_pipeline_engine = engine
            module.forward = engine.run
            if callable(getattr(module, "generate", None)) and not hasattr(type(module), "_pipelined_next_token"):
                # 🤗's GenerationMixin drives ``forward`` with cache / return_dict arguments and reads logits on every
                # rank: on a pipelined model that ends in an obscure shape error on the stages without the head
                module.generate = _generate_not_pipelined
            self._save_metadata(module, ctx)
        return module

    @torch.no_grad()
    def deparallelize(self) -> nn.Module:
        """Undo :meth:`parallelize`: every stage broadcasts its parameters over the PIPELINE group so that each rank
        holds the whole module again, and ``module.forward`` is the model's own forward (collective; unimplemented
        in the reference)."""
        import torch.distributed as dist

        from pipegoose_b200.distributed.parallel_mode import ParallelMode

        module, ctx = self.module, self.parallel_context
        if ctx.pipeline_parallel_size == 1 or not hasattr(module, "_pg_pipeline_stage"):
            return module
        group = ctx.get_group(ParallelMode.PIPELINE)
        ranks = ctx.get_ranks_in_group(ParallelMode.PIPELINE)
        stage_ids = {id(p) for p in module._pg_pipeline_stage.parameters()}
        mine = {n: tuple(p.shape) for n, p in module.named_parameters() if id(p) in stage_ids}
        owned = [None] * len(ranks)
        dist.all_gather_object(owned, mine, group=group)
        backend_dev = ctx.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        for name, p in module.named_parameters():
            owner = next(i for i, shapes in enumerate(owned) if name in shapes)  # tied table: the first stage
            shape = owned[owner][name]
            if id(p) in stage_ids and tuple(p.shape) == shape:
                buf = p.data.to(backend_dev)
            else:
                buf = torch.empty(shape, dtype=p.dtype, device=backend_dev)
            dist.broadcast(buf, src=ranks[owner], group=group)
            if not (id(p) in stage_ids and tuple(p.shape) == shape):
                p.data = buf.to(p.device if p.numel() > 0 else (ctx.device if backend_dev.type == "cuda" else "cpu"))
                p.requires_grad_(True)
        for attr in ("_pg_pipeline_stage", "_pg_pipeline_engine"):
            if hasattr(module, attr):
                delattr(module, attr)
        if module.__dict__.get("generate") is _generate_not_pipelined:
            del module.__dict__["generate"]
        if "forward" in module.__dict__:
            del module.__dict__["forward"]  # back to the class's forward
        return module


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
