"""Worker threads of the job runtime (parity: reference nn/pipeline_parallel/_worker.py:12-188).

Same roles as the reference — a selector moves jobs from the pending to the selected queue, workers
execute selected jobs, a watcher grows the pool up to ``max_workers`` when every worker is busy — but
every thread blocks on a queue / event instead of polling with sleeps or busy-spinning, the pool can be
destroyed, and a failing job is recorded (``WorkerManager.failed_jobs``) instead of killing its thread
silently."""
from __future__ import annotations

import threading
from queue import Empty, Queue
from typing import List, Optional

from pipegoose_b200.constants import PIPELINE_MAX_WORKERS, PIPELINE_MIN_WORKERS
from pipegoose_b200.nn.pipeline_parallel.queue import JobQueue

_STOP = object()


class Worker(threading.Thread):
    """Executes jobs taken from the selected-jobs queue."""

    def __init__(self, selected_jobs: Queue, finished_jobs: Optional[Queue] = None, failed: Optional[list] = None,
                 busy_event: Optional[threading.Event] = None, **kwargs):
        super().__init__(daemon=True, **kwargs)
        self._selected_jobs = selected_jobs
        self._finished_jobs = finished_jobs
        self._failed = failed if failed is not None else []
        self._busy_event = busy_event
        self._running = False
        self.lock = threading.Lock()

    @property
    def is_running(self) -> bool:
        return self._running

    def run(self):
        while True:
            job = self._selected_jobs.get()
            if job is _STOP:
                self._selected_jobs.task_done()
                return
            with self.lock:
                self._running = True
                if self._busy_event is not None:
                    self._busy_event.set()
                try:
                    job.compute()
                    if self._finished_jobs is not None:
                        self._finished_jobs.put(job)
                except BaseException as e:  # recorded; the pool stays alive
                    self._failed.append((job, e))
                finally:
                    self._running = False
                    self._selected_jobs.task_done()


class JobSelector(threading.Thread):
    """Moves jobs from the pending queue to the selected queue (FIFO)."""

    def __init__(self, pending_jobs: Queue, selected_jobs: Queue, **kwargs):
        super().__init__(daemon=True, **kwargs)
        self._pending_jobs = pending_jobs
        self._selected_jobs = selected_jobs

    def run(self):
        while True:
            job = self._pending_jobs.get()
            if job is _STOP:
                self._pending_jobs.task_done()
                return
            self._selected_jobs.put(job)   # counted as unfinished in `selected` BEFORE it leaves `pending`
            self._pending_jobs.task_done()


class WorkerPoolWatcher(threading.Thread):
    """Adds a worker when all current workers are busy and the pool is below ``max_workers``."""

    def __init__(self, worker_pool: List[Worker], min_workers: int, max_workers: int, spawn_worker, busy_event,
                 stop_event, **kwargs):
        super().__init__(daemon=True, **kwargs)
        self.worker_pool = worker_pool
        self.min_workers, self.max_workers = min_workers, max_workers
        self.spawn_worker = spawn_worker
        self._busy_event, self._stop_event = busy_event, stop_event

    def _num_working_workers(self) -> int:
        return sum(1 for w in self.worker_pool if w.is_running)

    def run(self):
        while not self._stop_event.is_set():
            self._busy_event.wait(timeout=0.5)  # a worker just became busy
            self._busy_event.clear()
            n = self._num_working_workers()
            if n == len(self.worker_pool) and len(self.worker_pool) < self.max_workers:
                self.spawn_worker()


class BaseWorkerManager:
    """Life cycle of a worker pool (parity: reference _worker.py:88-95)."""

    def spawn(self):
        raise NotImplementedError

    def destroy(self):
        raise NotImplementedError


class WorkerManager(BaseWorkerManager):
    def __init__(self, num_workers: int = PIPELINE_MIN_WORKERS, min_workers: int = PIPELINE_MIN_WORKERS,
                 max_workers: int = PIPELINE_MAX_WORKERS, pending_jobs: Queue = None, selected_jobs: Queue = None):
        assert min_workers <= num_workers <= max_workers or (min_workers <= max_workers and num_workers <= max_workers)
        self.num_workers, self.min_workers, self.max_workers = num_workers, min_workers, max_workers
        self._pending_jobs = pending_jobs if pending_jobs is not None else JobQueue.PENDING_JOBS
        self._selected_jobs = selected_jobs if selected_jobs is not None else JobQueue.SELECTED_JOBS
        self._finished_jobs = JobQueue.FINISHED_JOBS
        self._worker_pool: List[Worker] = []
        self.failed_jobs: list = []
        self._busy = threading.Event()
        self._stop = threading.Event()
        self._selector: Optional[JobSelector] = None
        self._watcher: Optional[WorkerPoolWatcher] = None

    @property
    def pending_jobs(self) -> Queue:
        return self._pending_jobs

    @property
    def selected_jobs(self) -> Queue:
        return self._selected_jobs

    @property
    def worker_pool(self) -> List[Worker]:
        return self._worker_pool

    def _spawn_a_worker(self):
        w = Worker(self._selected_jobs, self._finished_jobs, self.failed_jobs, self._busy)
        w.start()
        self._worker_pool.append(w)

    def spawn(self):
        for _ in range(self.num_workers):
            self._spawn_a_worker()
        self._selector = JobSelector(self._pending_jobs, self._selected_jobs)
        self._selector.start()
        self._watcher = WorkerPoolWatcher(self._worker_pool, self.min_workers, self.max_workers, self._spawn_a_worker,
                                          self._busy, self._stop)
        self._watcher.start()

    def wait_idle(self, timeout: float = 30.0) -> bool:
        """Block until both queues are drained and no worker is running."""
        import time

        # Queue.unfinished_tasks counts a job from put() until the consumer's task_done(): a job is never
        # "nowhere" between the selector and a worker, so this cannot report idle while a job is in flight
        end = time.monotonic() + timeout
        while time.monotonic() < end:
            if self._pending_jobs.unfinished_tasks == 0 and self._selected_jobs.unfinished_tasks == 0:
                return True
            time.sleep(0.002)
        return False

    def destroy(self):
        self._stop.set()
        self._busy.set()
        if self._selector is not None:
            self._pending_jobs.put(_STOP)
            self._selector.join(timeout=2)
        for _ in self._worker_pool:
            self._selected_jobs.put(_STOP)
        for w in self._worker_pool:
            w.join(timeout=2)
        self._worker_pool.clear()
        for q in (self._pending_jobs, self._selected_jobs):  # drop sentinels nobody consumed
            try:
                while True:
                    if q.get_nowait() is not _STOP:
                        pass
            except Empty:
                pass
