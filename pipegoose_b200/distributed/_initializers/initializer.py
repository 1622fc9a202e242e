"""Process-group initializers (parity: reference distributed/_initializers/*.py).

One generic implementation walks the rank sets produced by :class:`Topology`; the per-mode
classes only pick the mode.  Every rank must call ``dist.new_group`` for every group, in the
same order, which ``Topology.groups`` guarantees.
"""
from __future__ import annotations

from abc import ABC
from typing import List, TypedDict

import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.distributed.topology import Topology


class ProcessGroupResult(TypedDict):
    local_rank: int
    local_world_size: int
    process_group: dist.ProcessGroup
    ranks_in_group: List[int]
    parallel_mode: ParallelMode


class ProcessGroupInitializer(ABC):
    parallel_mode: ParallelMode = None

    def __init__(self, rank: int, world_size: int, tensor_parallel_size: int, pipeline_parallel_size: int, data_parallel_size: int):
        self.rank = rank
        self.world_size = world_size
        self.tensor_parallel_size = tensor_parallel_size
        self.pipeline_parallel_size = pipeline_parallel_size
        self.data_parallel_size = data_parallel_size
        self.topology = Topology(world_size, tensor_parallel_size, pipeline_parallel_size, data_parallel_size)

    def init_dist_group(self) -> ProcessGroupResult:
        mine = None
        for ranks in self.topology.groups(self.parallel_mode):
            group = dist.new_group(ranks=ranks)
            if self.rank in ranks:
                mine = ProcessGroupResult(
                    local_rank=ranks.index(self.rank),
                    local_world_size=len(ranks),
                    process_group=group,
                    ranks_in_group=ranks,
                    parallel_mode=self.parallel_mode,
                )
        assert mine is not None
        return mine
