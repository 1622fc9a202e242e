"""Per-process view of the static pipeline schedule (parity: reference
nn/pipeline_parallel/pipeline_context.py:22-162).  No threads or condition variables: the
schedule is a table, ``clock_idx`` only advances when the engine finishes a clock."""
from __future__ import annotations

from enum import Enum, auto
from typing import List

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.nn.pipeline_parallel._utils import get_partition_idx
from pipegoose_b200.nn.pipeline_parallel.scheduler import BaseScheduler
from pipegoose_b200.nn.pipeline_parallel.task import Task


class TrainingState(Enum):
    IDLE = auto()
    FORWARD = auto()
    BACKWARD = auto()
    FINISHED = auto()


class PipelineContext:
    _current: "PipelineContext" = None    # the most recently constructed one of this process

    def __init__(self, scheduler: BaseScheduler, parallel_context: ParallelContext):
        self.scheduler = scheduler
        self.parallel_context = parallel_context
        self._clock_idx = 0
        self._state = TrainingState.IDLE
        PipelineContext._current = self

    @staticmethod
    def get_context() -> "PipelineContext":
        """The process's pipeline context (parity: reference pipeline_context.py:36-42, a module global)."""
        return PipelineContext._current

    # ------------------------------------------------------------------ state
    @property
    def state(self) -> TrainingState:
        return self._state

    def forward(self):
        self._state = TrainingState.FORWARD

    def backward(self):
        self._state = TrainingState.BACKWARD

    def finish(self):
        self._state = TrainingState.FINISHED

    # ------------------------------------------------------------------ topology
    @property
    def partition_idx(self) -> int:
        return get_partition_idx(self.parallel_context)

    @property
    def num_microbatches(self) -> int:
        return self.scheduler.n_microbatches

    @property
    def is_first_stage(self) -> bool:
        return self.partition_idx == 0

    @property
    def is_last_stage(self) -> bool:
        return self.partition_idx == self.parallel_context.pipeline_parallel_size - 1

    def is_last_microbatch(self, microbatch_idx: int) -> bool:
        return microbatch_idx == self.num_microbatches - 1

    # ------------------------------------------------------------------ clock
    @property
    def clock_idx(self) -> int:
        return self._clock_idx

    def increase_a_clock_cycle(self):
        self._clock_idx += 1

    def reset_clock(self):
        self._clock_idx = 0

    # ------------------------------------------------------------------ schedule lookups
    @property
    def schedules(self) -> List[List[Task]]:
        return self.scheduler.get_schedules()

    @property
    def schedule(self) -> List[Task]:
        """This partition's tasks in the current clock cycle."""
        return self.get_schedule_from_partition(self._clock_idx, self.partition_idx)

    def get_schedule(self):
        """Iterate this partition's tasks clock by clock, advancing the clock."""
        for clock in range(len(self.schedules)):
            self._clock_idx = clock
            yield self.get_schedule_from_partition(clock, self.partition_idx)

    def get_schedule_from_partition(self, clock_idx: int, partition_idx: int) -> List[Task]:
        return [t for t in self.schedules[clock_idx] if t.partition_idx == partition_idx]

    def get_schedule_from_microbatch(self, clock_idx: int, microbatch_idx: int) -> List[Task]:
        return [t for t in self.schedules[clock_idx] if t.microbatch_idx == microbatch_idx]

    def get_next_schedule_from_microbatch(self, microbatch_idx: int) -> List[Task]:
        return self.get_schedule_from_microbatch(self._clock_idx + 1, microbatch_idx)
