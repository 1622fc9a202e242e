"""Owns the bucket distributors of a process (the reference's BucketManager is an empty class)."""
from __future__ import annotations

from typing import Dict

import torch.distributed as dist

from pipegoose_b200.constants import BUCKET_SIZE_MB
from pipegoose_b200.core.bucket.dist import BucketDistributor


class BucketManager:
    def __init__(self, parallel_context, bucket_size_mb: float = BUCKET_SIZE_MB):
        self.parallel_context = parallel_context
        self.bucket_size_mb = bucket_size_mb
        self._distributors: Dict = {}

    def distributor(self, op=dist.all_reduce) -> BucketDistributor:
        if op not in self._distributors:
            self._distributors[op] = BucketDistributor(op, self.bucket_size_mb, self.parallel_context)
        return self._distributors[op]

    def flush(self):
        for d in self._distributors.values():
            d.flush()
