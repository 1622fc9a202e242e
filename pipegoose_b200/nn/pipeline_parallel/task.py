"""One cell of a pipeline schedule (parity: reference nn/pipeline_parallel/task.py:6-10): which job type runs for
which micro-batch on which partition.  Hashable, so schedules can be checked for duplicates and used as dict keys."""
from dataclasses import dataclass
from typing import Tuple

from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType


@dataclass(frozen=True)
class Task:
    job_type: JobType
    microbatch_idx: int
    partition_idx: int

    @property
    def key(self) -> Tuple[int, int]:
        """``(microbatch_idx, partition_idx)``: the key of the activation stores and of the progress tracker."""
        return (self.microbatch_idx, self.partition_idx)

    @property
    def is_forward(self) -> bool:
        return self.job_type is JobType.FORWARD
