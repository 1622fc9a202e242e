import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.core.bucket.bucket import Bucket
from pipegoose_b200.core.bucket.dist import BucketDistributor
from pipegoose_b200.core.bucket.exception import BucketClosedError, BucketFullError
from pipegoose_b200.core.bucket.utils import mb_size_to_num_elements
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.testing.utils import init_parallel_context, spawn
from pipegoose_b200.utils.memory import get_tensor_storage_mem_loc


def test_bucket_aliases_added_tensors():
    bucket = Bucket(16, torch.float32)
    a, b = torch.arange(6.0).view(2, 3), torch.arange(4.0)
    assert bucket.available_size == 16 and len(bucket) == 0 and bucket.is_free
    bucket.add_tensor(a), bucket.add_tensor(b)
    assert len(bucket) == 2 and bucket.available_size == 6 and not bucket.is_full
    assert get_tensor_storage_mem_loc(a) == bucket.storage().data_ptr() == get_tensor_storage_mem_loc(b)
    assert a.tolist() == [[0, 1, 2], [3, 4, 5]]
    a.add_(1)  # in-place update of the tensor is an update of the bucket
    assert bucket.buffer()[:6].tolist() == [1, 2, 3, 4, 5, 6]
    with pytest.raises(BucketFullError):
        bucket.add_tensor(torch.zeros(7))
    bucket.close()
    assert bucket.is_closed
    with pytest.raises(BucketClosedError):
        bucket.add_tensor(torch.zeros(1))
    bucket.clear()
    assert not bucket.is_closed and bucket.available_size == 16 and len(bucket) == 0


def test_mb_size_to_num_elements():
    assert mb_size_to_num_elements(1, torch.float32) == 262144
    assert mb_size_to_num_elements(1, torch.bfloat16) == 524288
    assert mb_size_to_num_elements(25, torch.float16) == 25 * 524288
    with pytest.raises(ValueError):
        mb_size_to_num_elements(1, torch.complex64)


def run_distributor(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    d = BucketDistributor(dist.all_reduce, bucket_size_mb=0.001, parallel_context=ctx)  # 262 fp32 elements
    big = torch.full((1000,), float(rank + 1))      # larger than a bucket: reduced on its own
    smalls = [torch.full((100,), float(rank + 1) * (i + 1)) for i in range(5)]  # packed, several flushes
    d.execute(big, ParallelMode.DATA)
    for t in smalls:
        d.execute(t, ParallelMode.DATA)
    d.flush()
    total = sum(range(1, world_size + 1))
    assert torch.all(big == total)
    for i, t in enumerate(smalls):
        assert torch.all(t == total * (i + 1)), i
    ctx.destroy()


def test_bucket_distributor_packs_and_flushes():
    spawn(run_distributor, world_size=2)
