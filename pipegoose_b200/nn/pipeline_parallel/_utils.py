"""Pipeline helpers (parity: reference nn/pipeline_parallel/_utils.py:7-22)."""
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def get_partition_idx(parallel_context: ParallelContext) -> int:
    """Index of the pipeline stage this rank holds."""
    rank = parallel_context.get_global_rank()
    return parallel_context.get_ranks_in_group(ParallelMode.PIPELINE).index(rank)


def is_last_stage(parallel_context: ParallelContext) -> bool:
    return get_partition_idx(parallel_context) == parallel_context.pipeline_parallel_size - 1


def is_first_stage(parallel_context: ParallelContext) -> bool:
    return get_partition_idx(parallel_context) == 0


def sleep(seconds: float = 0.05):
    """Blocking sleep (parity: reference nn/pipeline_parallel/_utils.py:7-9; the runtime itself never polls)."""
    import time

    time.sleep(seconds)
