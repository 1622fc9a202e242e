"""Put jobs into a queue (parity: reference nn/pipeline_parallel/_job/register.py:6-17)."""
from queue import Queue

from pipegoose_b200.nn.pipeline_parallel._job.job import Job


class _JobRegister:
    def __init__(self, queue: Queue):
        self.queue = queue

    def registry(self, job: Job):
        assert isinstance(job, Job), f"job must be a Job, got {type(job)}"
        self.queue.put(job)


def add_job_to_queue(job: Job, queue: Queue):
    _JobRegister(queue).registry(job)
