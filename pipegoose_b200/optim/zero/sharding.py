"""Greedy optimizer-state sharding at parameter granularity (parity: reference
optim/zero/sharding.py:10-46): every parameter goes to the rank that currently holds the fewest
elements; each rank receives the same number of param groups (possibly with empty ``params``)."""
from __future__ import annotations

from typing import Dict, List

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class OptimizerStateSharding:
    def __init__(self, param_groups: List[Dict], parallel_context: ParallelContext, parallel_mode: ParallelMode):
        self.param_groups = param_groups
        self.parallel_context = parallel_context
        self.parallel_mode = parallel_mode

    def shard(self) -> List[List[Dict]]:
        world = self.parallel_context.get_world_size(self.parallel_mode)
        loads = [0] * world
        sharded: List[List[Dict]] = [[] for _ in range(world)]
        for group in self.param_groups:
            per_rank = [[] for _ in range(world)]
            for p in group["params"]:
                target = min(range(world), key=lambda r: (loads[r], r))
                per_rank[target].append(p)
                loads[target] += p.numel()
            for r in range(world):
                g = {k: v for k, v in group.items() if k != "params"}
                g["params"] = per_rank[r]
                sharded[r].append(g)
        return sharded
