"""A pre-allocated flat buffer that tensors are packed into (parity: reference
core/bucket/bucket.py:6-88).  ``add_tensor`` copies the tensor in and re-points ``tensor.data`` at
the bucket's storage so later in-place updates of the tensor are updates of the bucket."""
from __future__ import annotations

import torch

from pipegoose_b200.core.bucket.exception import BucketClosedError, BucketFullError


class Bucket:
    def __init__(self, size: int, dtype: torch.dtype, device=None):
        assert size > 0, "bucket size must be positive"
        self.size = size
        self.dtype = dtype
        self._buffer = torch.zeros(size, dtype=dtype, device=device)
        self._offset = 0
        self._is_closed = False
        self._num_tensors = 0

    @property
    def is_closed(self) -> bool:
        return self._is_closed

    @property
    def available_size(self) -> int:
        return self.size - self._offset

    @property
    def is_full(self) -> bool:
        return self._offset >= self.size

    @property
    def is_free(self) -> bool:
        return self._offset == 0

    def add_tensor(self, tensor: torch.Tensor) -> torch.Tensor:
        assert isinstance(tensor, torch.Tensor), "only tensors can be added to a bucket"
        assert tensor.dtype == self.dtype, "tensor dtype differs from the bucket dtype"
        if self._is_closed:
            raise BucketClosedError("the bucket is closed")
        n = tensor.numel()
        if n > self.available_size:
            raise BucketFullError("the bucket has not enough space for this tensor")
        view = self._buffer[self._offset:self._offset + n]
        view.copy_(tensor.detach().reshape(-1))
        tensor.data = view.view_as(tensor)  # alias the bucket storage
        self._offset += n
        self._num_tensors += 1
        return tensor

    def storage(self):
        return self._buffer.untyped_storage()

    def buffer(self) -> torch.Tensor:
        """The filled part of the bucket."""
        return self._buffer[:self._offset]

    def close(self):
        assert not self._is_closed, "the bucket is already closed"
        self._is_closed = True

    def clear(self):
        assert self._offset > 0, "the bucket is already empty"
        self._offset = 0
        self._num_tensors = 0
        self._is_closed = False
        self._buffer.zero_()

    def __len__(self) -> int:
        return self._num_tensors
