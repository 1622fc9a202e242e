"""Schedule entry (parity: reference nn/pipeline_parallel/task.py:6-10)."""
from dataclasses import dataclass

from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType


@dataclass(frozen=True)
class Task:
    """One unit of pipeline work: run ``job_type`` for a micro-batch on a partition."""

    job_type: JobType
    microbatch_idx: int
    partition_idx: int
