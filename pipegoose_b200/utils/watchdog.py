"""Failure detection (the reference has none: a rank that dies or a job that raises leaves every other rank blocked
in a collective forever — SURVEY §5.3).

Two small tools, both host-side and off the hot path:

* :class:`RankWatchdog` — every rank publishes a heartbeat counter in the job's c10d key-value store (the TCPStore
  that ``init_process_group`` created; no extra sockets, no collectives, nothing on the GPU) from a daemon thread,
  and checks its peers' counters.  A peer whose counter has not moved for ``timeout_s`` is reported through
  ``on_failure(dead_ranks)``; the default handler prints every thread's stack and sends the process a SIGINT, which
  turns a silent hang into a ``RankFailure`` (and, with ``abort_after_s``, ends a process whose main thread is stuck
  inside a CUDA / NCCL call so that the launcher can restart the job).
  With ``stall_timeout_s`` it also watches its OWN process for progress (:meth:`RankWatchdog.tick`): a wedged main
  thread — whose heartbeat thread would keep every peer happy forever — ends the process.
* :func:`step_deadline` — ``with step_deadline(120): train_step()`` dumps all thread stacks (``faulthandler``) when a
  step overruns, which is what one wants to see from a job that is stuck inside NCCL or a spinning flag wait.
"""
from __future__ import annotations

import faulthandler
import sys
import threading
import time
from contextlib import contextmanager
from typing import Callable, Dict, List, Optional

import torch.distributed as dist


STALL_EXIT_CODE = 75      # EX_TEMPFAIL: "try again" — what a launcher with --max-restarts does


class RankFailure(RuntimeError):
    """Raised in the main thread when peers stopped heart-beating."""

    def __init__(self, dead_ranks: List[int]):
        super().__init__(f"ranks {dead_ranks} stopped responding")
        self.dead_ranks = dead_ranks


class RankWatchdog:
    def __init__(self, parallel_context=None, timeout_s: float = 60.0, interval_s: float = 1.0,
                 on_failure: Optional[Callable[[List[int]], None]] = None, ranks: Optional[List[int]] = None,
                 store=None, tag: str = "watchdog", abort_after_s: Optional[float] = None,
                 stall_timeout_s: Optional[float] = None):
        """``ranks``: the global ranks to watch (default: the whole job).  ``store``: any c10d ``Store`` (default: the
        job's own).  ``stall_timeout_s``: the PROGRESS deadline — the owner calls :meth:`tick` whenever it makes progress
        (the ``Trainer``: after every micro-step, evaluation batch and checkpoint); when no tick arrives for this long
        the main thread is wedged (a sleeping callback, a spinning kernel, a collective whose peer is wedged — the
        heartbeat THREAD of such a process keeps beating, so no peer would ever call it dead): every thread's stack is
        dumped and the process ends with :data:`STALL_EXIT_CODE`, which takes the attempt down so that the launcher can
        restart it.  Must exceed the longest legitimate gap between ticks (first step, checkpoint writes)."""
        self.rank = parallel_context.get_global_rank() if parallel_context is not None else dist.get_rank()
        world = dist.get_world_size()
        self.ranks = [r for r in (ranks if ranks is not None else range(world)) if r != self.rank]
        base = store if store is not None else dist.distributed_c10d._get_default_store()
        self.store = dist.PrefixStore(f"pg_b200/{tag}", base)
        self.timeout_s, self.interval_s = float(timeout_s), float(interval_s)
        self.abort_after_s = abort_after_s  # default handler only: hard-exit if the main thread ignores the interrupt
        self.on_failure = on_failure or self._default_handler
        self.failed: List[int] = []
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._beat = 0
        self._pending_failure: Optional[RankFailure] = None
        self.stall_timeout_s = stall_timeout_s
        self._last_tick = time.monotonic()
        self.ticks = 0

    # ------------------------------------------------------------------ lifecycle
    def tick(self):
        """The owner made progress (see ``stall_timeout_s``)."""
        self.ticks += 1
        self._last_tick = time.monotonic()

    def start(self) -> "RankWatchdog":
        assert self._thread is None, "already started"
        self._last_tick = time.monotonic()
        self._publish()
        self._thread = threading.Thread(target=self._loop, name=f"pg-watchdog-{self.rank}", daemon=True)
        self._thread.start()
        return self

    def stop(self, announce: bool = True):
        """Stop watching.  ``announce``: tell the peers this is a clean exit, not a failure."""
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=self.interval_s * 4 + 1)
            self._thread = None
        if announce:
            try:
                self.store.set(f"bye/{self.rank}", "1")
            except Exception:
                pass

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
        return False

    def check(self):
        """Raise :class:`RankFailure` if the watchdog saw dead peers (call between steps when interrupting the main
        thread is not wanted: ``RankWatchdog(..., on_failure=lambda dead: None)``)."""
        if self.failed:
            raise RankFailure(list(self.failed))

    # ------------------------------------------------------------------ internals
    def _publish(self):
        self._beat += 1
        self.store.set(f"beat/{self.rank}", str(self._beat))

    def _read(self, key: str) -> Optional[str]:
        try:
            if not self.store.check([key]):
                return None
            return self.store.get(key).decode()
        except Exception:
            return None

    def _loop(self):
        last_value: Dict[int, Optional[str]] = {r: None for r in self.ranks}
        last_change: Dict[int, float] = {r: time.monotonic() for r in self.ranks}
        while not self._stop.wait(self.interval_s):
            try:
                self._publish()
            except Exception:
                # the store lives in rank 0's process: losing it means rank 0 is gone
                self._report([0] if self.rank != 0 else [])
                return
            now = time.monotonic()
            if self.stall_timeout_s is not None and now - self._last_tick > self.stall_timeout_s:
                self._stalled(now - self._last_tick)
                return
            dead = []
            for r in self.ranks:
                if r in self.failed:
                    continue
                if self._read(f"bye/{r}") is not None:
                    last_change[r] = now  # left on purpose
                    continue
                v = self._read(f"beat/{r}")
                if v != last_value[r]:
                    last_value[r], last_change[r] = v, now
                elif now - last_change[r] > self.timeout_s:
                    dead.append(r)
            if dead:
                self._report(dead)

    def _stalled(self, silent_for: float):
        """The main thread cannot be asked to raise — it is the one that is stuck.  Say why, then end the process."""
        import os

        sys.stderr.write(f"[pipegoose_b200 watchdog] rank {self.rank}: no progress for {silent_for:.0f} s "
                         f"(after {self.ticks} ticks), ending the process so that the launcher can restart the job\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
        os._exit(STALL_EXIT_CODE)

    def _report(self, dead: List[int]):
        new = [r for r in dead if r not in self.failed]
        if not new:
            return
        self.failed.extend(new)
        self.on_failure(list(self.failed))

    def _default_handler(self, dead: List[int]):
        sys.stderr.write(f"[pipegoose_b200 watchdog] rank {self.rank}: ranks {dead} stopped responding\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        import os
        import signal

        self._pending_failure = RankFailure(dead)
        # a real SIGINT (not ``_thread.interrupt_main``): it also wakes a main thread blocked in sleep / select / a
        # lock; callers convert the resulting KeyboardInterrupt with ``translate()``
        os.kill(os.getpid(), signal.SIGINT)
        if self.abort_after_s is not None:
            # a main thread stuck inside a CUDA / NCCL call never returns to the interpreter: give it a grace period,
            # then end the process so that the launcher (torchrun --max-restarts) can restart the job
            if not self._stop.wait(self.abort_after_s):
                sys.stderr.write(f"[pipegoose_b200 watchdog] rank {self.rank}: main thread did not react, exiting\n")
                sys.stderr.flush()
                os._exit(1)

    @contextmanager
    def translate(self):
        """``with watchdog.translate(): train()`` — the interrupt the default handler sends becomes a
        :class:`RankFailure` naming the dead ranks."""
        try:
            yield
        except KeyboardInterrupt:
            if self._pending_failure is not None:
                raise self._pending_failure from None
            raise


@contextmanager
def step_deadline(seconds: float, file=None, exit_on_timeout: bool = False):
    """Dump every thread's stack if the body runs longer than ``seconds`` (and optionally terminate the process so
    that the launcher can restart the job)."""
    faulthandler.dump_traceback_later(seconds, repeat=False, file=file or sys.stderr, exit=exit_on_timeout)
    try:
        yield
    finally:
        faulthandler.cancel_dump_traceback_later()
