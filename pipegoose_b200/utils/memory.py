import torch


def get_tensor_storage_mem_loc(tensor: torch.Tensor) -> int:
    """Address of the tensor's underlying storage (parity: reference utils/memory.py:4-6)."""
    return tensor.untyped_storage().data_ptr()
