"""Static pipeline schedules (parity: reference nn/pipeline_parallel/scheduler.py:35-115 for GPipe;
1F1B is new).  Schedules are pure functions of ``(n_microbatches, n_partitions)``: every rank
computes the same table, so no control-plane traffic is needed at run time (the reference
confirms every task to a master over RPC).

* ``get_schedules()`` — list of clock cycles, each a list of :class:`Task` that run concurrently.
* ``get_stage_order(partition_idx)`` — the ordered task list one stage executes.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum, auto
from typing import List

from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel.task import Task


class SchedulerType(Enum):
    GPIPE = auto()
    ONE_F_ONE_B = auto()


class BaseScheduler(ABC):
    def __init__(self, n_microbatches: int, n_partitions: int):
        assert n_microbatches > 0 and n_partitions > 0, "microbatches and partitions must be positive"
        self.n_microbatches = n_microbatches
        self.n_partitions = n_partitions

    @abstractmethod
    def get_schedules(self) -> List[List[Task]]:
        raise NotImplementedError

    @abstractmethod
    def get_stage_order(self, partition_idx: int) -> List[Task]:
        raise NotImplementedError

    def get_forward_schedules(self) -> List[List[Task]]:
        return [[t for t in clock if t.job_type is JobType.FORWARD] for clock in self.get_schedules()
                if any(t.job_type is JobType.FORWARD for t in clock)]

    def get_backward_schedules(self) -> List[List[Task]]:
        return [[t for t in clock if t.job_type is JobType.BACKWARD] for clock in self.get_schedules()
                if any(t.job_type is JobType.BACKWARD for t in clock)]

    # ------------------------------------------------------------------ timing model of a schedule
    def simulate(self, forward_cost: float = 1.0, backward_cost: float = 2.0, transfer_cost: float = 0.0):
        """Event simulation of the per-stage task orders under their data dependencies: every stage runs its tasks
        in order, a forward of micro-batch ``i`` on stage ``p`` starts once stage ``p - 1`` finished it (+ one transfer),
        a backward once stage ``p + 1`` finished its backward (the last stage: its own forward).

        Returns ``(makespan, busy, start)``: the length of the step, the busy time per stage, and ``start[task]``.
        This is what the static runtime executes (one task at a time per stage, batched p2p between neighbours), so it
        predicts the step time of a layout from a stage's measured forward / backward time."""
        n = self.n_partitions
        orders = [self.get_stage_order(p) for p in range(n)]
        cost = {JobType.FORWARD: forward_cost, JobType.BACKWARD: backward_cost}
        finish, start = {}, {}
        cursor, clock = [0] * n, [0.0] * n
        remaining = sum(len(o) for o in orders)
        while remaining:
            progressed = False
            for p in range(n):
                while cursor[p] < len(orders[p]):
                    t = orders[p][cursor[p]]
                    if t.job_type is JobType.FORWARD:
                        dep = None if p == 0 else Task(JobType.FORWARD, t.microbatch_idx, p - 1)
                    elif p == n - 1:
                        dep = Task(JobType.FORWARD, t.microbatch_idx, p)
                    else:
                        dep = Task(JobType.BACKWARD, t.microbatch_idx, p + 1)
                    if dep is not None and dep not in finish:
                        break
                    ready = 0.0 if dep is None else finish[dep] + (transfer_cost if dep.partition_idx != p else 0.0)
                    start[t] = max(clock[p], ready)
                    finish[t] = clock[p] = start[t] + cost[t.job_type]
                    cursor[p] += 1
                    remaining -= 1
                    progressed = True
            assert progressed, "schedule deadlocked"
        makespan = max(finish.values())
        busy = [sum(cost[t.job_type] for t in o) for o in orders]
        return makespan, busy, start

    def bubble_fraction(self, forward_cost: float = 1.0, backward_cost: float = 2.0, transfer_cost: float = 0.0) -> float:
        """Share of the step a stage spends idle, averaged over the stages: ``1 - busy / makespan``.  With equal stages
        and free transfers this is ``(n - 1) / (m + n - 1)`` for both GPipe and 1F1B."""
        makespan, busy, _ = self.simulate(forward_cost, backward_cost, transfer_cost)
        return 1.0 - sum(busy) / (len(busy) * makespan)

    def peak_live_microbatches(self, partition_idx: int) -> int:
        """Most micro-batches whose activations stage ``partition_idx`` holds at once (forward done, backward not yet):
        ``m`` for GPipe, ``min(n - p, m)`` for 1F1B — the reason to prefer 1F1B at equal bubble."""
        live = peak = 0
        for t in self.get_stage_order(partition_idx):
            live += 1 if t.job_type is JobType.FORWARD else -1
            peak = max(peak, live)
        return peak

    @property
    def total_clock_cycles(self) -> int:
        return len(self.get_schedules())

    @property
    def total_forward_clock_cycles(self) -> int:
        return len(self.get_forward_schedules())

    @property
    def total_backward_clock_cycles(self) -> int:
        return len(self.get_backward_schedules())


class GPipeScheduler(BaseScheduler):
    """All forwards, then all backwards: at forward clock ``c`` partition ``p`` runs micro-batch ``c - p``."""

    def _forward_clocks(self) -> List[List[Task]]:
        m, n = self.n_microbatches, self.n_partitions
        clocks = []
        for c in range(m + n - 1):
            clocks.append([Task(JobType.FORWARD, c - p, p) for p in range(n) if 0 <= c - p < m])
        return clocks

    def get_schedules(self) -> List[List[Task]]:
        fwd = self._forward_clocks()
        bwd = [[Task(JobType.BACKWARD, t.microbatch_idx, t.partition_idx) for t in clock] for clock in reversed(fwd)]
        return fwd + bwd

    def get_stage_order(self, partition_idx: int) -> List[Task]:
        return [t for clock in self.get_schedules() for t in clock if t.partition_idx == partition_idx]


class OneFOneBScheduler(BaseScheduler):
    """PipeDream-flush / 1F1B: stage ``p`` runs ``n - 1 - p`` warm-up forwards, then alternates one
    forward with one backward, then drains the remaining backwards.  Same bubble as GPipe but at
    most ``n - p`` micro-batches of activations alive per stage."""

    def get_stage_order(self, partition_idx: int) -> List[Task]:
        m, n, p = self.n_microbatches, self.n_partitions, partition_idx
        warmup = min(n - 1 - p, m)
        order: List[Task] = [Task(JobType.FORWARD, i, p) for i in range(warmup)]
        f, b = warmup, 0
        while f < m:
            order.append(Task(JobType.FORWARD, f, p))
            f += 1
            order.append(Task(JobType.BACKWARD, b, p))
            b += 1
        while b < m:
            order.append(Task(JobType.BACKWARD, b, p))
            b += 1
        return order

    def get_schedules(self) -> List[List[Task]]:
        """Clock-cycle view obtained by simulating the per-stage orders under their data dependencies."""
        n = self.n_partitions
        orders = [self.get_stage_order(p) for p in range(n)]
        cursor = [0] * n
        done = set()
        clocks: List[List[Task]] = []
        total = sum(len(o) for o in orders)
        while len(done) < total:
            ready = []
            for p in range(n):
                if cursor[p] >= len(orders[p]):
                    continue
                t = orders[p][cursor[p]]
                if t.job_type is JobType.FORWARD:
                    dep = None if p == 0 else Task(JobType.FORWARD, t.microbatch_idx, p - 1)
                else:
                    dep = Task(JobType.FORWARD, t.microbatch_idx, p) if p == n - 1 else Task(JobType.BACKWARD, t.microbatch_idx, p + 1)
                if dep is None or dep in done:
                    ready.append(t)
            assert ready, "1F1B schedule deadlocked"
            for t in ready:
                cursor[t.partition_idx] += 1
            done.update(ready)
            clocks.append(ready)
        return clocks


def get_scheduler(scheduler_type: SchedulerType):
    mapping = {SchedulerType.GPIPE: GPipeScheduler, SchedulerType.ONE_F_ONE_B: OneFOneBScheduler}
    return mapping[scheduler_type]
