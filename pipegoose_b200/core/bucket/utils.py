"""Bucket sizing helpers (parity: reference core/bucket/utils.py:4-29; bf16 added)."""
import torch

_BYTES = {
    torch.float64: 8, torch.float32: 4, torch.float16: 2, torch.bfloat16: 2,
    torch.int64: 8, torch.int32: 4, torch.int16: 2, torch.int8: 1, torch.uint8: 1, torch.bool: 1,
}


def mb_size_to_num_elements(mb: float, dtype: torch.dtype) -> int:
    """Number of ``dtype`` elements that fit in ``mb`` megabytes."""
    if dtype not in _BYTES:
        raise ValueError(f"unsupported dtype: {dtype}")
    return int(mb * 1024 * 1024 // _BYTES[dtype])


def get_memory_address_of_tensor_storage(tensor: torch.Tensor) -> int:
    return tensor.untyped_storage().data_ptr()
