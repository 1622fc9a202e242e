"""A pre-allocated flat buffer that tensors are packed into back to back (parity: reference core/bucket/bucket.py:6-88).

Packing a tensor copies it to the cursor position and re-points ``tensor.data`` at that span of the bucket, so the
tensor and the bucket are the same memory from then on: a collective over ``bucket.buffer()`` moves every packed
tensor, and in-place updates of a tensor are updates of the bucket (the property the reference's tests pin with
``data_ptr`` comparisons).
"""
from __future__ import annotations

from typing import Iterator, List, Tuple

import torch

from pipegoose_b200.core.bucket.exception import BucketClosedError, BucketFullError


class Bucket:
    __slots__ = ("size", "dtype", "_flat", "_cursor", "_sealed", "_spans")

    def __init__(self, size: int, dtype: torch.dtype, device=None):
        assert size > 0, "bucket size must be positive"
        self.size, self.dtype = size, dtype
        self._flat = torch.zeros(size, dtype=dtype, device=device)
        self._cursor = 0                           # first free element
        self._sealed = False
        self._spans: List[Tuple[int, torch.Size]] = []   # (offset, shape) of every packed tensor, in packing order

    # ---- state -----------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._spans)

    def __iter__(self) -> Iterator[torch.Tensor]:
        """Views of the packed tensors, in packing order."""
        for offset, shape in self._spans:
            yield self._flat[offset:offset + shape.numel()].view(shape)

    available_size = property(lambda self: self.size - self._cursor)
    is_free = property(lambda self: self._cursor == 0)
    is_full = property(lambda self: self._cursor >= self.size)
    is_closed = property(lambda self: self._sealed)

    # ---- packing ---------------------------------------------------------------------------------------------
    def add_tensor(self, tensor: torch.Tensor) -> torch.Tensor:
        assert isinstance(tensor, torch.Tensor), "only tensors can be added to a bucket"
        assert tensor.dtype == self.dtype, "tensor dtype differs from the bucket dtype"
        if self._sealed:
            raise BucketClosedError("the bucket is closed")
        count = tensor.numel()
        if count > self.available_size:
            raise BucketFullError("the bucket has not enough space for this tensor")
        span = self._flat.narrow(0, self._cursor, count)
        span.copy_(tensor.detach().reshape(-1))
        tensor.data = span.view(tensor.shape)      # from here on the tensor lives inside the bucket
        self._spans.append((self._cursor, tensor.shape))
        self._cursor += count
        return tensor

    def buffer(self) -> torch.Tensor:
        """The filled part of the bucket (what a collective should move)."""
        return self._flat.narrow(0, 0, self._cursor)

    def storage(self):
        return self._flat.untyped_storage()

    # ---- life cycle ------------------------------------------------------------------------------------------
    def close(self):
        assert not self._sealed, "the bucket is already closed"
        self._sealed = True

    def clear(self):
        assert self._cursor > 0, "the bucket is already empty"
        self._flat.zero_()
        self._spans.clear()
        self._cursor, self._sealed = 0, False
