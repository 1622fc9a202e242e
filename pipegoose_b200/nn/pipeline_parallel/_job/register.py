"""Admission of jobs into a queue (the reference's nn/pipeline_parallel/_job/register.py:6-17 wraps ``queue.put``).

Here the register is the one place that decides whether something may enter a worker queue: it must be a
:class:`Job` that has not started yet — a job that is already executing, done or failed is never run twice, which
is what a re-queued callback or a retried package would otherwise cause.
"""
from __future__ import annotations

from queue import Queue

from pipegoose_b200.nn.pipeline_parallel._job.job import Job, JobStatus


class _JobRegister:
    __slots__ = ("queue",)

    def __init__(self, queue: Queue):
        self.queue = queue

    def admissible(self, job) -> bool:
        return isinstance(job, Job) and job.status is JobStatus.PENDING

    def registry(self, job: Job) -> Job:
        assert isinstance(job, Job), f"only Job objects can be queued, got {type(job).__name__}"
        assert self.admissible(job), f"job {job.key} is {job.status.name}: only pending jobs can be queued"
        self.queue.put(job)
        return job


def add_job_to_queue(job: Job, queue: Queue) -> Job:
    return _JobRegister(queue).registry(job)
