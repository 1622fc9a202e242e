"""Which pipeline stage am I (parity: reference nn/pipeline_parallel/_utils.py:7-22).

A rank's stage index is its position inside its PIPELINE group; with the library's rank layout (pp outermost) that is
``global_rank // (world // pp)``, but the group list is the source of truth so custom layouts keep working.
"""
from __future__ import annotations

import time

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def get_partition_idx(parallel_context: ParallelContext) -> int:
    stages = parallel_context.get_ranks_in_group(ParallelMode.PIPELINE)
    return stages.index(parallel_context.get_global_rank())


def is_first_stage(parallel_context: ParallelContext) -> bool:
    return get_partition_idx(parallel_context) == 0


def is_last_stage(parallel_context: ParallelContext) -> bool:
    return get_partition_idx(parallel_context) + 1 == parallel_context.pipeline_parallel_size


def sleep(timeout: float = 0.05):
    """The reference's polling interval helper (:7-9).  Kept for API parity; nothing in this runtime polls."""
    time.sleep(timeout)
