"""Stage-to-stage transport of :class:`Package` objects (parity: reference nn/pipeline_parallel/_comm.py:9-41).

The reference pushes pickled packages with ``rpc.rpc_sync`` into a process-global ``RECV_QUEUE`` on the
receiver.  Here a package is two point-to-point messages on the PIPELINE group (NCCL on GPUs, gloo on
CPU): a fixed-size int64 header carrying the metadata and the typed tensor payload; the receiver calls
:func:`recv_package` (static schedules know who sends next) which also files the package in
``RECV_QUEUE`` for consumers written against the queue interface."""
from __future__ import annotations

from queue import Queue

import torch
import torch.distributed as dist

from pipegoose_b200.distributed._p2p import _P2P, _comm_device
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._package import Metadata, Package, TrainingMetadata

RECV_QUEUE: Queue = Queue()
_HEADER = 7  # microbatch, partition, job_type, is_training, is_grad_enabled, src, dst


def _local(rank_global: int, ctx: ParallelContext) -> int:
    return ctx.get_ranks_in_group(ParallelMode.PIPELINE).index(rank_global)


def send_package(package: Package, parallel_context: ParallelContext):
    """Send ``package`` to ``package.metadata.dst`` (a global rank of this rank's PIPELINE group)."""
    assert isinstance(package, Package), f"expected a Package, got {type(package)}"
    m = package.metadata
    group = parallel_context.get_group(ParallelMode.PIPELINE)
    dev = _comm_device(parallel_context, group)
    header = torch.tensor([m.microbatch_idx, m.partition_idx, m.job_type.value, int(m.training.is_training),
                           int(m.training.is_grad_enabled), m.src, m.dst], dtype=torch.long)
    dist.send(header.to(dev), dst=m.dst, group=group)
    _P2P().send(package.data, _local(m.dst, parallel_context), parallel_context, ParallelMode.PIPELINE)


def recv_package(src: int, parallel_context: ParallelContext, enqueue: bool = True) -> Package:
    """Receive the next package from global rank ``src``."""
    group = parallel_context.get_group(ParallelMode.PIPELINE)
    dev = _comm_device(parallel_context, group)
    header = torch.zeros(_HEADER, dtype=torch.long, device=dev)
    dist.recv(header, src=src, group=group)
    h = header.cpu().tolist()
    data = _P2P().recv(_local(src, parallel_context), parallel_context, ParallelMode.PIPELINE)
    package = Package(data, Metadata(h[0], h[1], JobType(h[2]), TrainingMetadata(bool(h[3]), bool(h[4])), h[5], h[6]))
    if enqueue:
        RECV_QUEUE.put(package)
    return package
