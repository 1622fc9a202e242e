"""Which data-parallel rank keeps the optimizer state of which parameter (parity: reference optim/zero/sharding.py:10-46).

Longest-processing-time greedy at parameter granularity: parameters are handed, in order, to the rank with the smallest
load so far (a min-heap of ``(elements, rank)``), which keeps the shards within one parameter of each other.  Every rank
gets the same number of param groups — possibly with an empty ``params`` list — so that group indices and
hyper-parameters line up across ranks.  (The fused ZeRO-1 path does not use this: it slices the flat buffer evenly,
``FusedAdam.set_bucket_shards``.)
"""
from __future__ import annotations

import heapq
from typing import Dict, List

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class OptimizerStateSharding:
    def __init__(self, param_groups: List[Dict], parallel_context: ParallelContext, parallel_mode: ParallelMode):
        self.param_groups = param_groups
        self.parallel_context = parallel_context
        self.parallel_mode = parallel_mode

    def shard(self) -> List[List[Dict]]:
        n_ranks = self.parallel_context.get_world_size(self.parallel_mode)
        heap = [(0, rank) for rank in range(n_ranks)]          # (elements owned so far, rank)
        plan: List[List[Dict]] = [[] for _ in range(n_ranks)]
        for group in self.param_groups:
            hyper = {key: value for key, value in group.items() if key != "params"}
            owned: List[List] = [[] for _ in range(n_ranks)]
            for param in group["params"]:
                load, rank = heapq.heappop(heap)
                owned[rank].append(param)
                heapq.heappush(heap, (load + param.numel(), rank))
            for rank in range(n_ranks):
                plan[rank].append({**hyper, "params": owned[rank]})
        return plan
