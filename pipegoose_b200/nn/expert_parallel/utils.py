"""Where the experts of a layer live (parity: reference nn/expert_parallel/utils.py:5-8).

The experts of one MoE layer are dealt out in contiguous blocks over the TENSOR (== EXPERT) group: rank ``r`` owns
global experts ``[r * n_local, (r + 1) * n_local)``.
"""
from __future__ import annotations

from typing import Tuple

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def local_expert_range(num_experts: int, parallel_context: ParallelContext) -> Tuple[int, int]:
    """``(first, last + 1)`` global expert indices owned by this rank."""
    group = parallel_context.get_world_size(ParallelMode.TENSOR)
    per_rank, left_over = divmod(num_experts, group)
    assert left_over == 0, f"{num_experts} experts cannot be dealt evenly over a tensor group of {group}"
    first = parallel_context.get_local_rank(ParallelMode.TENSOR) * per_rank
    return first, first + per_rank


def get_num_local_experts(num_experts: int, parallel_context: ParallelContext) -> int:
    first, end = local_expert_range(num_experts, parallel_context)
    return end - first
