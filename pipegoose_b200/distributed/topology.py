"""Rank <-> (pp, dp, tp) coordinate arithmetic.

The layout is fixed by the reference's group initializers
(distributed/_initializers/initialize_{tensor,pipeline,data}.py): tensor-parallel ranks are
contiguous (fastest axis), data-parallel is the middle axis and pipeline stages are the
outermost axis, i.e. ``global = pp * (dp * tp) + dp * tp_size + tp``.  With that layout a
tensor-parallel group always sits on neighbouring GPUs of one NVSwitch domain, which is what
the fused GEMM+collective kernels want.

Everything here is pure Python so it can be unit-tested without a process group.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

from pipegoose_b200.distributed.parallel_mode import ParallelMode


@dataclass(frozen=True)
class Coord:
    pp: int
    dp: int
    tp: int


class Topology:
    def __init__(self, world_size: int, tensor_parallel_size: int, pipeline_parallel_size: int, data_parallel_size: int):
        assert tensor_parallel_size * pipeline_parallel_size * data_parallel_size == world_size, (
            "tensor_parallel_size * pipeline_parallel_size * data_parallel_size must equal the world size: "
            f"{tensor_parallel_size} * {pipeline_parallel_size} * {data_parallel_size} != {world_size}"
        )
        self.world_size = world_size
        self.tp = tensor_parallel_size
        self.pp = pipeline_parallel_size
        self.dp = data_parallel_size

    # ------------------------------------------------------------------ coordinates
    def coord(self, rank: int) -> Coord:
        per_stage = self.dp * self.tp
        return Coord(pp=rank // per_stage, dp=(rank % per_stage) // self.tp, tp=rank % self.tp)

    def rank_of(self, pp: int, dp: int, tp: int) -> int:
        return pp * (self.dp * self.tp) + dp * self.tp + tp

    # ------------------------------------------------------------------ groups
    def groups(self, mode: ParallelMode) -> List[List[int]]:
        """All rank sets of ``mode``, in creation order (identical on every rank)."""
        if mode is ParallelMode.GLOBAL:
            return [list(range(self.world_size))]
        if mode in (ParallelMode.TENSOR, ParallelMode.EXPERT):
            return [
                [self.rank_of(p, d, t) for t in range(self.tp)] for p in range(self.pp) for d in range(self.dp)
            ]
        if mode in (ParallelMode.DATA, ParallelMode.EXPERT_DATA):
            return [
                [self.rank_of(p, d, t) for d in range(self.dp)] for p in range(self.pp) for t in range(self.tp)
            ]
        if mode is ParallelMode.PIPELINE:
            return [
                [self.rank_of(p, d, t) for p in range(self.pp)] for d in range(self.dp) for t in range(self.tp)
            ]
        raise ValueError(f"unknown parallel mode {mode}")

    def group_of(self, rank: int, mode: ParallelMode) -> List[int]:
        for g in self.groups(mode):
            if rank in g:
                return g
        raise ValueError(f"rank {rank} is in no {mode} group")

    def local_rank(self, rank: int, mode: ParallelMode) -> int:
        return self.group_of(rank, mode).index(rank)

    def describe(self, rank: int) -> Dict[ParallelMode, int]:
        return {m: self.local_rank(rank, m) for m in ParallelMode}
