"""NVLink peer-mapped ("symmetric") workspaces.

Every rank of a process group allocates the same-sized device buffer, exports it with CUDA IPC and
maps every peer's buffer, so kernels can ``ld/st.global`` (and ``cp.async.bulk``) directly on
peer memory through NVSwitch.  The first ``SIGNAL_BYTES`` of each buffer are 32-bit flags /
counters used by the in-kernel protocols (system-scope release/acquire); the rest is data.

This is the substrate for the fused GEMM+collective kernels (``ops/comm.py``) and the fused
data-parallel / ZeRO-1 kernels; NCCL stays the bootstrap and the fallback.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

SIGNAL_BYTES = 64 * 1024
# u32 indices inside the signal area
SIG_BARRIER = 0        # [0,16): two barrier phases x 8 peers; [16,18): grid counters (see csrc/comm.cu)
SIG_AG_READY = 64      # [64,72): "rank src's shard for all-gather epoch e is readable"
SIG_RS_ARRIVE = 128    # [128,136): tiles pushed into my reduce-scatter staging by rank src
SIG_CHUNK_CTR = 192    # [192,200): local: comm CTAs that finished copying chunk c


_NODE_CACHE = {}


def peers_share_a_node(parallel_context, parallel_mode) -> bool:
    """True when every rank of the group runs on this host and the group fits one NVSwitch domain — the precondition of
    CUDA-IPC peer mapping.  Groups that span hosts (e.g. the DATA group of a multi-node job) keep the NCCL paths.
    Collective over the group (one ``all_gather_object`` of the host identities, cached per group)."""
    import os
    import socket

    from pipegoose_b200.constants import MAX_NVLINK_PEERS

    ranks = tuple(parallel_context.get_ranks_in_group(parallel_mode))
    if ranks in _NODE_CACHE:
        return _NODE_CACHE[ranks]
    if len(ranks) == 1:
        same = True
    else:
        boot = ""
        try:
            boot = open("/proc/sys/kernel/random/boot_id").read().strip()
        except OSError:
            pass
        me = (socket.gethostname(), boot, os.environ.get("PIPEGOOSE_B200_FAKE_NODE", ""))
        everyone = [None] * len(ranks)
        dist.all_gather_object(everyone, me, group=parallel_context.get_group(parallel_mode))
        same = all(e == everyone[0] for e in everyone) and len(ranks) <= MAX_NVLINK_PEERS
    _NODE_CACHE[ranks] = same
    return same


class SymmetricWorkspace:
    def __init__(self, parallel_context, parallel_mode, nbytes: int):
        from pipegoose_b200.ops import native

        self._n = native()
        self.ctx = parallel_context
        self.mode = parallel_mode
        self.group = parallel_context.get_group(parallel_mode)
        self.world = parallel_context.get_world_size(parallel_mode)
        self.rank = parallel_context.get_local_rank(parallel_mode)
        self.nbytes = int(nbytes)
        self.device_index = torch.cuda.current_device()
        total = SIGNAL_BYTES + ((self.nbytes + 1023) // 1024) * 1024
        self._total = total
        self._local_ptr, handle = self._n.symm_alloc(total)
        handles: List[bytes] = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        self.peer_ptrs: List[int] = []
        for r, h in enumerate(handles):
            self.peer_ptrs.append(self._local_ptr if r == self.rank else self._n.symm_open(h))
        self._bytes = self._n.tensor_from_ptr(self._local_ptr, total, self.device_index)
        self._closed = False
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------ addressing
    def data_ptr(self, rank: int, offset: int = 0) -> int:
        return self.peer_ptrs[rank] + SIGNAL_BYTES + offset

    def sig_ptr(self, rank: int, index: int) -> int:
        return self.peer_ptrs[rank] + 4 * index

    def local_tensor(self, offset: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + nbytes <= self.nbytes, "symmetric workspace overflow"
        start = SIGNAL_BYTES + offset
        return self._bytes[start:start + nbytes].view(dtype).view(*shape)

    def signals(self) -> torch.Tensor:
        return self._bytes[:SIGNAL_BYTES].view(torch.int32)

    def close(self):
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        except Exception:
            pass
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self._n.symm_close(p)
        self._bytes = None
        self._n.symm_free(self._local_ptr)
