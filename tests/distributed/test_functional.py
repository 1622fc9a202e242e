import pytest
import torch

from pipegoose_b200.distributed import functional as F
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.testing.utils import init_parallel_context, spawn

MODES = [ParallelMode.GLOBAL, ParallelMode.TENSOR, ParallelMode.PIPELINE, ParallelMode.DATA]


def run_collectives(rank, world_size, port, tp, pp, dp):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    for mode in MODES:
        n = ctx.get_world_size(mode)
        lr = ctx.get_local_rank(mode)
        ranks = ctx.get_ranks_in_group(mode)
        # scatter: local slice, no comm
        x = torch.arange(n * 2 * 3, dtype=torch.float32).reshape(3, n * 2)
        assert torch.equal(F.scatter(x, dim=-1, parallel_context=ctx, parallel_mode=mode), x[:, lr * 2:(lr + 1) * 2])
        # all_reduce
        t = torch.tensor([float(rank)])
        F.all_reduce(t, parallel_context=ctx, parallel_mode=mode)
        assert t.item() == float(sum(ranks))
        # all_gather on dim 0 and last dim, and 0-d
        g = F.all_gather(torch.full((2, 3), float(rank)), dim=0, parallel_context=ctx, parallel_mode=mode)
        assert g.shape == (2 * n, 3) and torch.equal(g[::2, 0], torch.tensor(ranks, dtype=torch.float32))
        g = F.all_gather(torch.full((2, 3), float(rank)), dim=-1, parallel_context=ctx, parallel_mode=mode)
        assert g.shape == (2, 3 * n) and torch.equal(g[0, ::3], torch.tensor(ranks, dtype=torch.float32))
        g = F.all_gather(torch.tensor(float(rank)), parallel_context=ctx, parallel_mode=mode)
        if n > 1:
            assert g.tolist() == [float(r) for r in ranks]
        # broadcast / reduce address ranks by their GLOBAL rank, like torch.distributed and the reference's tests
        b = torch.tensor([float(rank)])
        F.broadcast(b, src=ranks[-1], parallel_context=ctx, parallel_mode=mode)
        assert b.item() == float(ranks[-1])
        r = torch.tensor([1.0])
        F.reduce(r, dst=ranks[0], parallel_context=ctx, parallel_mode=mode)
        if lr == 0:
            assert r.item() == float(n)
        # reduce_scatter
        rs_in = torch.arange(n * 4, dtype=torch.float32).reshape(n * 2, 2) + rank
        rs = F.reduce_scatter(rs_in.clone(), dim=0, parallel_context=ctx, parallel_mode=mode)
        base = torch.arange(n * 4, dtype=torch.float32).reshape(n * 2, 2)
        expect = (base * n + sum(ranks))[lr * 2:(lr + 1) * 2] if n > 1 else rs_in
        assert torch.equal(rs, expect)
        # all_to_all (equal splits): row r of rank s goes to rank r
        a2a_in = torch.full((n, 2), float(lr)) + torch.arange(n, dtype=torch.float32).unsqueeze(1) * 10
        a2a = F.all_to_all(a2a_in, parallel_context=ctx, parallel_mode=mode)
        assert torch.equal(a2a[:, 0], torch.arange(n, dtype=torch.float32) + lr * 10)
        F.barrier(ctx, mode)
    ctx.destroy()


@pytest.mark.parametrize("world_size,tp,pp,dp", [(1, 1, 1, 1), (8, 2, 2, 2)])
def test_functional_collectives(world_size, tp, pp, dp):
    spawn(run_collectives, world_size=world_size, tp=tp, pp=pp, dp=dp)


def run_p2p(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, 1, world_size, 1)
    for dtype, rg in [(torch.float32, True), (torch.bfloat16, False), (torch.int64, False)]:
        data = (torch.arange(12).reshape(3, 4)).to(dtype)
        if rg:
            data.requires_grad_(True)
        F.send(data, src=0, dst=1, parallel_context=ctx, parallel_mode=ParallelMode.PIPELINE)
        got = F.recv(src=0, dst=1, parallel_context=ctx, parallel_mode=ParallelMode.PIPELINE)
        if rank == 1:
            assert got.dtype == dtype and got.shape == (3, 4) and got.requires_grad == rg
            assert torch.equal(got.detach(), data.detach())
        else:
            assert got is None
    ctx.destroy()


def test_p2p():
    spawn(run_p2p, world_size=2)
