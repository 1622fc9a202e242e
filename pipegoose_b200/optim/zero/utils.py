"""Flatten / unflatten helpers (parity: reference optim/zero/utils.py:7-26)."""
from typing import List

import torch


def delete_tensor_from_memory(tensor: torch.Tensor):
    """Release a tensor's storage (the python object may still be referenced elsewhere)."""
    tensor.data = torch.empty(0, dtype=tensor.dtype, device=tensor.device)
    del tensor


def flatten_a_list_tensor(list: List[torch.Tensor]) -> torch.Tensor:  # noqa: A002 (the reference's keyword)
    return torch.cat([t.reshape(-1) for t in list]) if len(list) > 0 else torch.empty(0)


def copy_flatten_tensor_to_unflatten_tensors(flat: torch.Tensor, tensors: List[torch.Tensor]):
    offset = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[offset:offset + n].view_as(t))
        offset += n
    assert offset == flat.numel(), "flat tensor and tensor list sizes differ"
