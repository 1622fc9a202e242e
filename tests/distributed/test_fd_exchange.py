"""File descriptors travel between the ranks of a node over AF_UNIX sockets (the transport of the VMM / multicast
symmetric workspace, distributed/symmetric.py) — checked here with descriptors of temporary files."""
import os
import tempfile

import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.distributed.symmetric import exchange_fds
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def run_exchange(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    group = ctx.get_group(ParallelMode.TENSOR)
    def anon_file(payload: bytes) -> int:
        f = tempfile.TemporaryFile()
        f.write(payload)
        f.flush()
        return os.dup(f.fileno())

    mine = anon_file(f"hello from {rank}".encode())
    extra = anon_file(b"extra") if rank == 0 else -1   # rank 0 offers one more (the multicast object's role)
    fds = exchange_fds(group, rank, world_size, [mine, extra])
    for peer in range(world_size):
        if peer == rank:
            continue
        assert os.pread(fds[peer][0], 64, 0).decode() == f"hello from {peer}"
        assert len(fds[peer]) == (2 if peer == 0 else 1)
        if peer == 0:
            assert os.pread(fds[0][1], 64, 0) == b"extra"
    dist.barrier(group=group)
    ctx.destroy()


def test_exchange_fds_between_ranks():
    spawn(run_exchange, world_size=3)
