"""3-D composition TensorParallel -> PipelineParallel -> DataParallel -> DistributedOptimizer(FusedAdam) with the
canonical training loop (forward, zero_grad, backward, step) on 8 gloo processes: the loss trajectory must
follow the single-process model (BASELINE.json config #5 in miniature: TP2 x PP2 x DP2 + ZeRO-1, 1F1B)."""
import copy

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=4, n_head=4)
STEPS = 3


def run(rank, world_size, port, tp, pp, dp, n_mb, state, ids, ref_losses, runtime="static", gpipe=False):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType

    model = PipelineParallel(model, num_microbatches=n_mb, parallel_context=ctx, runtime=runtime,
                             scheduler_type=SchedulerType.GPIPE if gpipe else SchedulerType.ONE_F_ONE_B).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    losses = []
    for _ in range(STEPS):
        out = model(local, labels=local)
        optim.zero_grad()       # after forward, as in the reference's README loop
        out.loss.backward()
        optim.step()
        losses.append(out.loss.item())
    t = torch.tensor(losses)
    if not ctx.is_last_rank(ParallelMode.PIPELINE):
        t.zero_()
    dist.all_reduce(t)
    mean = (t / (world_size // pp)).tolist()
    for a, b in zip(mean, ref_losses):
        assert abs(a - b) < 2e-3, (mean, ref_losses)
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,dp,variant", [(2, 2, 2, "1f1b"), (1, 2, 2, "1f1b"), (2, 2, 1, "1f1b"), (1, 4, 2, "1f1b"),
                                              (2, 2, 2, "gpipe"), (2, 2, 2, "jobs")])
def test_3d_training_follows_single_process(tp, pp, dp, variant):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (8, 8))
    n_mb = 2
    opt = FusedAdam(model.parameters(), lr=1e-2)
    chunks = [mb for rep in ids.chunk(dp) for mb in rep.chunk(n_mb)]
    ref_losses = []
    for _ in range(STEPS):
        opt.zero_grad()
        total = 0.0
        for mb in chunks:
            loss = model(mb, labels=mb).loss / len(chunks)
            loss.backward()
            total += loss.item()
        opt.step()
        ref_losses.append(total)
    assert ref_losses[-1] < ref_losses[0]
    spawn(run, world_size=tp * pp * dp, tp=tp, pp=pp, dp=dp, n_mb=n_mb, state=state, ids=ids, ref_losses=ref_losses,
          runtime="jobs" if variant == "jobs" else "static", gpipe=variant == "gpipe")


def run_pp_dp_stock_optimizer(rank, world_size, port, state, ids, ref_losses):
    """PP x DP with a plain torch optimizer and the canonical loop: the schedule's gradients survive zero_grad()."""
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 2)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-1)
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    losses = []
    for _ in range(STEPS):
        out = model(local, labels=local)
        optim.zero_grad()
        out.loss.backward()
        optim.step()
        losses.append(out.loss.item())
    t = torch.tensor(losses)
    if not ctx.is_last_rank(ParallelMode.PIPELINE):
        t.zero_()
    dist.all_reduce(t)
    mean = (t / 2).tolist()
    for a, b in zip(mean, ref_losses):
        assert abs(a - b) < 2e-3, (mean, ref_losses)
    ctx.destroy()


def test_pipeline_x_data_parallel_with_a_stock_optimizer():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (8, 8))
    opt = torch.optim.SGD(model.parameters(), lr=1e-1)
    chunks = [mb for rep in ids.chunk(2) for mb in rep.chunk(2)]
    ref_losses = []
    for _ in range(STEPS):
        opt.zero_grad()
        total = 0.0
        for mb in chunks:
            loss = model(mb, labels=mb).loss / len(chunks)
            loss.backward()
            total += loss.item()
        opt.step()
        ref_losses.append(total)
    assert ref_losses[-1] < ref_losses[0]
    spawn(run_pp_dp_stock_optimizer, world_size=4, state=state, ids=ids, ref_losses=ref_losses)


def run_3d_deparallelize(rank, world_size, port, state, ids, ref_logits):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 2)
    model = BloomForCausalLM(BloomConfig(vocab_size=101, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    tp = TensorParallel(model, ctx)
    model = tp.parallelize()
    pp = PipelineParallel(model, num_microbatches=2, parallel_context=ctx)
    model = pp.parallelize()
    dp = DataParallel(model, ctx)
    model = dp.parallelize()
    model(ids, labels=ids)                      # a scheduled step on the sharded model
    for wrapper in (dp, pp, tp):                # undo in reverse order: every rank ends with the whole model
        model = wrapper.deparallelize()
    logits = model(ids).logits
    assert logits.shape == ref_logits.shape and torch.allclose(logits, ref_logits, atol=1e-4)
    consolidated = model.state_dict()
    for k, v in state.items():
        assert consolidated[k].shape == v.shape and torch.allclose(consolidated[k], v, atol=1e-6), k
    ctx.destroy()


def test_3d_deparallelize_gives_back_the_whole_model():
    """TP x PP x DP (vocabulary 101: padded for TP) -> deparallelize all three -> the original state dict and logits."""
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=101, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 101, (4, 8))
    spawn(run_3d_deparallelize, world_size=8, state=copy.deepcopy(model.state_dict()), ids=ids,
          ref_logits=model(ids).logits.detach())
