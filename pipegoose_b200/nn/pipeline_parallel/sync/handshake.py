"""Cross-rank progress tracking and group handshakes without RPC (parity: reference
nn/pipeline_parallel/sync/handshake.py:33-265).

The reference keeps a table ``{clock: {task: done}}`` consistent by RPC-ing every confirmation to a master
which re-broadcasts the table to every rank (with two ``time.sleep(0.1)`` per task).  Here the table lives
in the job's c10d key-value store (the TCPStore ``init_process_group`` already created): ``confirm`` is one
``set`` plus one atomic ``add``; whoever completes a clock cycle publishes a "done" key; every rank derives
the current clock and the table from the store when it looks, or blocks on the "done" key with
:meth:`ProgressTracker.wait_for_clock`.  No sleeps, no polling threads, no master-side bookkeeping."""
from __future__ import annotations

import pickle
from abc import ABC, abstractmethod
from typing import Dict, List, Optional

import torch.distributed as dist

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.pipeline_parallel.sync.callback import Callback

Progress = Dict[int, Dict[object, bool]]
_PROGRESS_TRACKER: Optional["ProgressTracker"] = None
_INSTANCES = 0


def set_progress_tracker(progress_tracker):
    global _PROGRESS_TRACKER
    _PROGRESS_TRACKER = progress_tracker


def get_progress_tracker() -> Optional["ProgressTracker"]:
    return _PROGRESS_TRACKER


def _store_for(parallel_context: ParallelContext, parallel_mode: ParallelMode, tag: str):
    """A namespaced view of the default store, unique per (tag, group)."""
    base = dist.distributed_c10d._get_default_store()
    ranks = parallel_context.get_ranks_in_group(parallel_mode)
    return dist.PrefixStore(f"pg_b200/{tag}/{parallel_mode.name}/{ranks[0]}-{ranks[-1]}-{len(ranks)}", base)


class Handshake(ABC):
    # class-level defaults so the attributes can be read off the class like in the reference (handshake.py:34-38, where
    # they are process-wide class state); here every handshake instance keeps its own
    master_rank: Optional[int] = None
    callbacks: List[Callback] = ()
    parallel_context: ParallelContext = None
    parallel_mode: ParallelMode = None

    def __init__(self, master_rank: int, callbacks: List[Callback] = (), parallel_context: ParallelContext = None,
                 parallel_mode: ParallelMode = ParallelMode.GLOBAL):
        self.master_rank = master_rank
        self.callbacks = list(callbacks)
        self.parallel_context = parallel_context
        self.parallel_mode = parallel_mode

    @abstractmethod
    def initiate(self, *args, **kwargs):
        ...

    @abstractmethod
    def confirm(self, *args, **kwargs):
        ...

    @abstractmethod
    def is_initiated(self) -> bool:
        ...

    @abstractmethod
    def is_confirmed(self, clock_idx: int = None) -> bool:
        ...

    @abstractmethod
    def is_all_confirmed(self, clock_idx: int = None) -> bool:
        ...


class ProgressTracker(Handshake):
    """Tracks which (clock cycle, task) pairs have completed across a process group."""

    def __init__(self, master_rank, callbacks=(), parallel_context=None, parallel_mode=ParallelMode.GLOBAL):
        super().__init__(master_rank, callbacks, parallel_context, parallel_mode)
        global _INSTANCES
        # every rank constructs its trackers in the same order -> the same generation id
        self._gen = _INSTANCES
        _INSTANCES += 1
        self._store = _store_for(parallel_context, parallel_mode, f"progress{self._gen}")
        self._round = 0
        self._progress: Optional[Progress] = None
        self._clock_idx = 0
        set_progress_tracker(self)

    # ------------------------------------------------------------------ keys
    def _k(self, name: str) -> str:
        return f"r{self._round}/{name}"

    def _load(self) -> bool:
        if self._progress is None:
            if not self._store.check([self._k("init")]):
                return False
            self._progress = pickle.loads(self._store.get(self._k("init")))
            self._clock_idx = 0
        return True

    # ------------------------------------------------------------------ protocol
    def initiate(self, progress: Progress):
        """Publish the table of tasks per clock cycle (master only; other ranks pick it up lazily)."""
        if self.parallel_context.get_global_rank() == self._global_master():
            # the previous table must be complete on every rank before a new one replaces it: a slower stage may
            # still be confirming tasks of the old schedule
            if self._load() and len(self._progress) > 0:
                self.wait_for_clock(len(self._progress) - 1)
            # a re-initiation (e.g. the backward schedule after the forward one) starts a new round
            while self._store.check([self._k("init")]):
                self._round += 1
            self._store.set(self._k("init"), pickle.dumps(progress))
            self._store.set("round", str(self._round))
        self._progress = None
        self._clock_idx = 0

    def _global_master(self) -> int:
        return self.parallel_context.get_global_rank_from_local_rank(self.master_rank, self.parallel_mode) \
            if self.parallel_mode != ParallelMode.GLOBAL else self.master_rank

    def _sync_round(self):
        if self._store.check(["round"]):
            r = int(self._store.get("round"))
            if r > self._round:  # rounds only move forward (the key is written after the table it announces)
                self._round, self._progress, self._clock_idx = r, None, 0

    def is_initiated(self) -> bool:
        self._sync_round()
        return self._load()

    def wait_initiated(self, round_idx: int, timeout_s: float = 60.0):
        """Block until the master published the table of round ``round_idx`` (0 for the first ``initiate``)."""
        import datetime

        self._store.wait([f"r{round_idx}/init"], datetime.timedelta(seconds=timeout_s))
        if self._round != round_idx:
            self._round, self._progress, self._clock_idx = round_idx, None, 0
        assert self._load()

    def confirm(self, task) -> None:
        """Mark ``task`` as done, for every rank.  A task that belongs to a later clock cycle than the current one
        (a stage running ahead of the others) first waits until the cycles before it are complete."""
        assert self.is_initiated(), "the progress tracker was not initiated"
        self._refresh()
        clock = self._clock_idx
        if clock >= len(self._progress) or task not in self._progress[clock]:
            later = [c for c in range(clock + 1, len(self._progress)) if task in self._progress[c]]
            assert later, f"task {task!r} is not part of clock cycle {clock} or any later one"
            self.wait_for_clock(later[0] - 1)
            clock = self._clock_idx
        assert task in self._progress[clock], f"task {task!r} is not part of clock cycle {clock}"
        self._store.set(self._k(f"c{clock}/{task!r}"), "1")
        n = self._store.add(self._k(f"n{clock}"), 1)
        if n == len(self._progress[clock]):
            self._store.set(self._k(f"done{clock}"), "1")
        self._refresh()

    def _refresh(self):
        """Advance the local clock over every completed cycle (firing callbacks) and fold in the
        confirmations of the current cycle."""
        if not self._load():
            return
        n_clocks = len(self._progress)
        while self._clock_idx < n_clocks and self._store.check([self._k(f"done{self._clock_idx}")]):
            for task in self._progress[self._clock_idx]:
                self._progress[self._clock_idx][task] = True
            self._clock_idx += 1
            for cb in sorted(self.callbacks, key=lambda c: c.order):
                cb.after_new_clock_cycle(self._progress, self._clock_idx)
        if self._clock_idx < n_clocks:
            cur = self._progress[self._clock_idx]
            for task in cur:
                if not cur[task] and self._store.check([self._k(f"c{self._clock_idx}/{task!r}")]):
                    cur[task] = True

    def wait_for_clock(self, clock_idx: int, timeout_s: float = 60.0):
        """Block until clock cycle ``clock_idx`` is complete on every rank (no polling)."""
        assert self.is_initiated()
        import datetime

        self._store.wait([self._k(f"done{clock_idx}")], datetime.timedelta(seconds=timeout_s))
        self._refresh()

    # ------------------------------------------------------------------ queries
    @property
    def clock_idx(self) -> int:
        self._refresh()
        return self._clock_idx

    @property
    def progress(self) -> Progress:
        self._refresh()
        return self._progress

    def is_confirmed(self, task, clock_idx: int) -> bool:
        self._refresh()
        return bool(self._progress[clock_idx][task])

    def is_all_confirmed(self, clock_idx: int) -> bool:
        self._refresh()
        return all(self._progress[clock_idx].values())


class ParallelGroupHandshake(Handshake):
    """Store-based rendezvous of one parallel group: every rank ``confirm``s, ``barrier`` returns when all have."""

    def __init__(self, parallel_context, parallel_mode, master_rank: int = 0, callbacks=()):
        super().__init__(master_rank, callbacks, parallel_context, parallel_mode)
        self._store = _store_for(parallel_context, parallel_mode, "handshake")
        self._world = parallel_context.get_world_size(parallel_mode)
        self._rank = parallel_context.get_local_rank(parallel_mode)
        self._epoch = 0

    def initiate(self):
        if self._rank == self.master_rank:
            self._store.set(f"e{self._epoch}/init", "1")

    def is_initiated(self) -> bool:
        return self._store.check([f"e{self._epoch}/init"])

    def confirm(self):
        self._store.set(f"e{self._epoch}/r{self._rank}", "1")
        if self._store.add(f"e{self._epoch}/n", 1) == self._world:
            self._store.set(f"e{self._epoch}/all", "1")

    def is_confirmed(self, rank: Optional[int] = None) -> bool:
        return self._store.check([f"e{self._epoch}/r{self._rank if rank is None else rank}"])

    def is_all_confirmed(self) -> bool:
        return self._store.check([f"e{self._epoch}/all"])

    def barrier(self, timeout_s: float = 60.0):
        import datetime

        if not self.is_confirmed():
            self.confirm()
        self._store.wait([f"e{self._epoch}/all"], datetime.timedelta(seconds=timeout_s))
        self._epoch += 1
