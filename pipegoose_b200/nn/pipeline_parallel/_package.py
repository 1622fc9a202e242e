"""What travels between pipeline stages in the job runtime: a tensor plus routing metadata
(parity: reference nn/pipeline_parallel/_package.py:8-37)."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any

from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType


@dataclass
class TrainingMetadata:
    is_training: bool
    is_grad_enabled: bool


@dataclass
class Metadata:
    """Routing information of a package."""

    microbatch_idx: int   # which micro-batch produced it
    partition_idx: int    # which pipeline partition must consume it
    job_type: JobType     # FORWARD: an activation, BACKWARD: a gradient
    training: TrainingMetadata
    src: int              # global rank that produced the data
    dst: int              # global rank that consumes it

    def key(self):
        return (self.microbatch_idx, self.partition_idx)


class Package:
    """A tensor (or tuple/dict of tensors) with its :class:`Metadata`."""

    def __init__(self, data: Any, metadata: Metadata):
        self.data = data
        self.metadata = metadata

    def clone_metadata(self, **changes) -> Metadata:
        return replace(self.metadata, **changes)

    def __repr__(self):
        m = self.metadata
        return (f"Package({m.job_type.name}, microbatch={m.microbatch_idx}, partition={m.partition_idx}, "
                f"{m.src}->{m.dst})")
