"""Storage-level tensor helpers (parity: reference utils/memory.py:4-6, used by the bucket tests to prove aliasing)."""
import torch


def get_tensor_storage_mem_loc(tensor: torch.Tensor) -> int:
    """Address of the first byte of the storage backing ``tensor`` (not of the view's first element)."""
    return tensor.untyped_storage().data_ptr()


def shares_storage(a: torch.Tensor, b: torch.Tensor) -> bool:
    """True when two tensors are views into the same allocation (e.g. a parameter re-pointed into a flat buffer)."""
    return get_tensor_storage_mem_loc(a) == get_tensor_storage_mem_loc(b)


def view_offset_bytes(tensor: torch.Tensor) -> int:
    """Byte offset of the view inside its storage."""
    return tensor.storage_offset() * tensor.element_size()
