"""DataParallel (bucketed flat-gradient reducer) and the ZeRO-1 DistributedOptimizer, fused and
generic paths, against single-process training on the full batch."""
import copy

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.optim.zero.sharding import OptimizerStateSharding
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)


def _reference(state, ids, steps, lr):
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        loss = model(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return {k: v.clone() for k, v in model.state_dict().items()}, losses


def run_dp(rank, world_size, port, dp, fused, state, ids, ref_state, ref_losses, lr):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, dp)
    torch.manual_seed(100 + rank)  # replicas start from DIFFERENT weights: DataParallel must broadcast rank 0's
    model = BloomForCausalLM(BloomConfig(**CFG))
    if rank == 0:
        model.load_state_dict(state)
    model = DataParallel(model, ctx, bucket_size_mb=0.02).parallelize()
    if fused:
        optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=lr), ctx)
    else:
        optim = DistributedOptimizer(torch.optim.Adam(model.parameters(), lr=lr), ctx)
    r = ctx.get_local_rank(ParallelMode.DATA)
    local = ids.chunk(dp)[r]
    for step in range(len(ref_losses)):
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
    assert len(model._pg_grad_reducer.buckets) > 1
    got = model.state_dict()
    for k, v in ref_state.items():
        assert torch.allclose(got[k], v, atol=2e-4), k
    if fused:
        # ZeRO-1: this rank only keeps optimizer state for ~1/dp of the parameters
        assert optim.optim.master.numel() * dp == optim.optim.flat.numel
    ctx.destroy()


@pytest.mark.parametrize("fused", [True, False])
def test_data_parallel_with_zero1_matches_full_batch_training(fused):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    ref_state, ref_losses = _reference(state, ids, steps=3, lr=1e-2)
    spawn(run_dp, world_size=2, dp=2, fused=fused, state=state, ids=ids, ref_state=ref_state, ref_losses=ref_losses, lr=1e-2)


def run_no_sync(rank, world_size, port, state, ids):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = DataParallel(model, ctx).parallelize()
    local = ids.chunk(2)[rank]
    model(local, labels=local).loss  # builds the reducer
    model._flat_state.zero_grad()
    with model.no_sync():
        model(local, labels=local).loss.backward()
    g_local = model._flat_state.flat_grad.clone()
    model._flat_state.zero_grad()
    model(local, labels=local).loss.backward()
    g_sync = model._flat_state.flat_grad.clone()
    other = [torch.zeros_like(g_local) for _ in range(2)]
    import torch.distributed as dist

    dist.all_gather(other, g_local)
    assert torch.allclose(g_sync, (other[0] + other[1]) / 2, atol=1e-6)
    assert not torch.allclose(g_local, g_sync)
    ctx.destroy()


def test_no_sync_skips_reduction():
    torch.manual_seed(0)
    state = BloomForCausalLM(BloomConfig(**CFG)).state_dict()
    spawn(run_no_sync, world_size=2, state=state, ids=torch.randint(0, 96, (4, 8)))


def run_sharding(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 3, 7, 5, 1, 8)]
    groups = [{"params": params[:4], "lr": 0.1}, {"params": params[4:], "lr": 0.2}]
    sharded = OptimizerStateSharding(groups, ctx, ParallelMode.DATA).shard()
    assert len(sharded) == world_size and all(len(s) == 2 for s in sharded)
    seen = [id(p) for s in sharded for g in s for p in g["params"]]
    assert sorted(seen) == sorted(id(p) for p in params)  # every parameter exactly once
    assert sharded[0][0]["lr"] == 0.1 and sharded[1][1]["lr"] == 0.2
    loads = [sum(p.numel() for g in s for p in g["params"]) for s in sharded]
    assert max(loads) - min(loads) <= 10
    ctx.destroy()


def test_optimizer_state_sharding():
    spawn(run_sharding, world_size=2)


def test_flatten_helpers():
    from pipegoose_b200.optim.zero.utils import copy_flatten_tensor_to_unflatten_tensors, flatten_a_list_tensor

    ts = [torch.arange(6.0).view(2, 3), torch.arange(4.0)]
    flat = flatten_a_list_tensor(ts)
    assert flat.tolist() == [0, 1, 2, 3, 4, 5, 0, 1, 2, 3]
    outs = [torch.zeros(2, 3), torch.zeros(4)]
    copy_flatten_tensor_to_unflatten_tensors(flat, outs)
    assert torch.equal(outs[0], ts[0]) and torch.equal(outs[1], ts[1])


def run_dp_deparallelize(rank, world_size, port, tp, state, ids, ref_grads):
    from pipegoose_b200.nn import TensorParallel

    ctx = init_parallel_context(rank, world_size, port, tp, 1, 2)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    if tp > 1:
        model = TensorParallel(model, ctx).parallelize()
    wrapper = DataParallel(model, ctx)
    model = wrapper.parallelize()
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    model(local, labels=local).loss.backward()           # a reduced step: the reducer is built and used
    model = wrapper.deparallelize()
    assert not hasattr(model, "_pg_grad_reducer")
    assert hasattr(model, "no_sync") == (tp > 1)           # handed back to the tensor-parallel partial-gradient sync
    model._flat_state.zero_grad(lazy=False)
    model(local, labels=local).loss.backward()           # must still run, now WITHOUT the data-parallel average
    for name in ("transformer.ln_f.weight", "transformer.h.0.input_layernorm.bias", "transformer.h.1.self_attention.dense.bias"):
        p = dict(model.named_parameters())[name]
        want = ref_grads[ctx.get_local_rank(ParallelMode.DATA)][name]
        assert torch.allclose(p.main_grad, want, atol=2e-5), name   # local-batch gradient, summed over the TENSOR group
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_data_parallel_deparallelize_detaches_the_reducer(tp):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    ref_grads = []
    for local in ids.chunk(2):
        model.zero_grad()
        model(local, labels=local).loss.backward()
        ref_grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
    spawn(run_dp_deparallelize, world_size=2 * tp, tp=tp, state=state, ids=ids, ref_grads=ref_grads)
