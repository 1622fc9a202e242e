import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.testing.utils import init_parallel_context, spawn

MODES = [ParallelMode.GLOBAL, ParallelMode.TENSOR, ParallelMode.PIPELINE, ParallelMode.DATA, ParallelMode.EXPERT_DATA]


def run_context(rank, world_size, port, tp, pp, dp):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    assert ParallelContext.get_context() is ctx
    assert ctx.tensor_parallel_size == tp and ctx.pipeline_parallel_size == pp and ctx.data_parallel_size == dp
    assert ctx.get_global_rank() == rank
    assert ctx.get_world_size(ParallelMode.GLOBAL) == world_size
    assert ctx.get_world_size(ParallelMode.TENSOR) == tp
    assert ctx.get_world_size(ParallelMode.PIPELINE) == pp
    assert ctx.get_world_size(ParallelMode.DATA) == dp
    for mode in MODES:
        assert ctx.is_initialized(mode)
        ranks = ctx.get_ranks_in_group(mode)
        lr = ctx.get_local_rank(mode)
        assert ranks[lr] == rank
        assert isinstance(ctx.get_group(mode), dist.ProcessGroup)
        assert ctx.get_global_rank_from_local_rank(lr, mode) == rank
        nxt, prv = ctx.get_next_global_rank(mode), ctx.get_prev_global_rank(mode)
        assert nxt == ranks[(lr + 1) % len(ranks)]
        assert prv == ranks[(lr - 1) % len(ranks)]
        assert ctx.is_first_rank(mode) == (lr == 0)
        assert ctx.is_last_rank(mode) == (lr == len(ranks) - 1)
        assert ctx.get_next_local_rank(lr, mode) == (lr + 1) % len(ranks)
        assert ctx.get_prev_local_rank(lr, mode) == (lr - 1) % len(ranks)
    # rank = pp * (dp*tp) + dp * tp + tp_rank
    expect = ctx.get_local_rank(ParallelMode.PIPELINE) * dp * tp + ctx.get_local_rank(ParallelMode.DATA) * tp + ctx.get_local_rank(ParallelMode.TENSOR)
    assert expect == rank
    key = (
        (ParallelMode.GLOBAL, rank),
        (ParallelMode.TENSOR, ctx.get_local_rank(ParallelMode.TENSOR)),
        (ParallelMode.PIPELINE, ctx.get_local_rank(ParallelMode.PIPELINE)),
        (ParallelMode.DATA, ctx.get_local_rank(ParallelMode.DATA)),
    )
    assert ctx.ranks2device(key) == rank
    assert ctx.get_worker_name(rank) == f"RPC_GLOBAL_WORKER_{rank}"
    # a real collective on every group
    for mode in MODES:
        t = torch.ones(1)
        dist.all_reduce(t, group=ctx.get_group(mode))
        assert t.item() == ctx.get_world_size(mode)
    ctx.destroy()
    for mode in MODES:
        assert not ctx.is_initialized(mode)
    assert ParallelContext.get_context() is None


@pytest.mark.parametrize("world_size,tp,pp,dp", [(1, 1, 1, 1), (8, 2, 2, 2), (4, 2, 1, 2), (16, 2, 4, 2)])
def test_parallel_context(world_size, tp, pp, dp):
    spawn(run_context, world_size=world_size, tp=tp, pp=pp, dp=dp)


def run_from_torch(rank, world_size, port):
    import os

    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_WORLD_SIZE=str(world_size),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ctx = ParallelContext.from_torch(tensor_parallel_size=2, pipeline_parallel_size=1, data_parallel_size=1)
    assert ctx.get_world_size(ParallelMode.TENSOR) == 2
    assert ctx.get_local_rank(ParallelMode.TENSOR) == rank
    ctx.destroy()


def test_from_torch():
    spawn(run_from_torch, world_size=2)


def test_bad_sizes_rejected():
    with pytest.raises(AssertionError):
        ParallelContext(rank=0, local_rank=0, world_size=4, local_world_size=4, host="127.0.0.1", port=1, seed=1,
                        backend="gloo", tensor_parallel_size=3, pipeline_parallel_size=1, data_parallel_size=1)


def _rpc_echo(x):
    return x * 2


def run_rpc(rank, world_size, port):
    """Opt-in RPC agents under the reference's worker names (parity: reference tests/distributed/test_rpc.py)."""
    from torch.distributed import rpc

    from pipegoose_b200.distributed.parallel_context import ParallelContext

    ctx = ParallelContext(rank=rank, local_rank=rank, world_size=world_size, local_world_size=world_size,
                          host="127.0.0.1", port=port, seed=69, backend="gloo", tensor_parallel_size=1,
                          pipeline_parallel_size=world_size, data_parallel_size=1, enable_rpc=True)
    peer = ctx.get_worker_name((rank + 1) % world_size)
    assert peer == f"RPC_GLOBAL_WORKER_{(rank + 1) % world_size}"
    assert rpc.rpc_sync(peer, _rpc_echo, args=(torch.tensor([rank + 1.0]),)).item() == 2.0 * (rank + 1)
    fut = rpc.rpc_async(peer, _rpc_echo, args=(torch.tensor([3.0]),))
    assert fut.wait().item() == 6.0
    ctx.destroy()  # shuts the agent down
    assert not ctx._rpc_started


def test_optional_rpc_workers():
    spawn(run_rpc, world_size=2)
