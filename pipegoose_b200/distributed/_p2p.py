"""Typed tensor send/recv between two ranks of a group (parity: reference distributed/_p2p.py).

Protocol: ONE fixed-size int64 header (dtype id, requires_grad, ndim, up to 8 dims) followed by
the payload, i.e. two messages instead of the reference's four.  When both sides already know
the shape (static pipeline schedules) use :func:`send_static` / :func:`recv_static`, which move
only the payload.  The receive buffer is allocated on the context's device (the reference
always allocated on the CPU).
"""
from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode

ID_TO_DTYPE = [
    torch.bfloat16,
    torch.float16,
    torch.float32,
    torch.float64,
    torch.uint8,
    torch.int8,
    torch.int16,
    torch.int32,
    torch.int64,
    torch.bool,
]
DTYPE_TO_ID = {dtype: idx for idx, dtype in enumerate(ID_TO_DTYPE)}

_MAX_DIMS = 8
_HEADER_LEN = 3 + _MAX_DIMS


def _comm_device(parallel_context, group):
    return parallel_context.device if dist.get_backend(group) == "nccl" else torch.device("cpu")


class _P2P:
    def send(self, data: Any, dst: int, parallel_context, parallel_mode: ParallelMode):
        if not isinstance(data, torch.Tensor):
            raise NotImplementedError(f"P2P send only supports torch.Tensor, got {type(data)}")
        group = parallel_context.get_group(parallel_mode)
        dst_global = parallel_context.get_global_rank_from_local_rank(dst, parallel_mode)
        dev = _comm_device(parallel_context, group)
        assert data.dim() <= _MAX_DIMS
        header = torch.zeros(_HEADER_LEN, dtype=torch.long)
        header[0] = DTYPE_TO_ID[data.dtype]
        header[1] = int(data.requires_grad)
        header[2] = data.dim()
        for i, s in enumerate(data.shape):
            header[3 + i] = s
        dist.send(header.to(dev), dst=dst_global, group=group)
        dist.send(data.detach().contiguous().to(dev), dst=dst_global, group=group)

    def recv(self, src: int, parallel_context, parallel_mode: ParallelMode) -> torch.Tensor:
        group = parallel_context.get_group(parallel_mode)
        src_global = parallel_context.get_global_rank_from_local_rank(src, parallel_mode)
        dev = _comm_device(parallel_context, group)
        header = torch.zeros(_HEADER_LEN, dtype=torch.long, device=dev)
        dist.recv(header, src=src_global, group=group)
        header = header.cpu().tolist()
        dtype = ID_TO_DTYPE[header[0]]
        shape = tuple(header[3:3 + header[2]])
        data = torch.empty(shape, dtype=dtype, device=dev)
        dist.recv(data, src=src_global, group=group)
        if header[1] and data.is_floating_point():
            data.requires_grad_(True)
        return data


def send_static(tensor: torch.Tensor, dst_global: int, group=None):
    """Payload-only send: the peer knows shape/dtype from the static schedule."""
    return dist.isend(tensor.contiguous(), dst=dst_global, group=group)


def recv_static(buffer: torch.Tensor, src_global: int, group=None):
    return dist.irecv(buffer, src=src_global, group=group)
