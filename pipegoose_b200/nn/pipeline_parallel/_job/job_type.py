"""Direction of a pipeline job (parity: reference nn/pipeline_parallel/_job/job_type.py:4-6).  The integer values
travel in package headers (``_comm.py``), so they are fixed rather than ``auto()``."""
from enum import IntEnum


class JobType(IntEnum):
    FORWARD = 1
    BACKWARD = 2

    @property
    def opposite(self) -> "JobType":
        """The job type that undoes / follows this one in a schedule (forward <-> backward)."""
        return JobType.BACKWARD if self is JobType.FORWARD else JobType.FORWARD

    @property
    def is_forward(self) -> bool:
        return self is JobType.FORWARD
