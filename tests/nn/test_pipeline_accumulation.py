"""Pipeline parallelism beyond one schedule per optimizer step: gradient accumulation over several schedules (against a
hand-written single-process loop, fused and stock optimizer) and loss evaluation under ``torch.no_grad()``."""
import copy

import pytest
import torch

from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn
from pipegoose_b200.trainer import Trainer

CFG = dict(vocab_size=96, hidden_size=32, n_layer=4, n_head=4)


def run_pp_accumulation(rank, world_size, port, state, data, ref_state, fused):
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    names = {id(p): n for n, p in m.named_parameters()}
    m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
    o = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5), ctx)
    Trainer(m, data, optim=o, parallel_context=ctx, grad_accum_steps=2).fit()
    for p in m._pg_pipeline_stage.parameters():
        assert torch.allclose(p.detach(), ref_state[names[id(p)]], atol=3e-5), names[id(p)]
    ctx.destroy()
@pytest.mark.parametrize("fused", [True, False])
def test_gradient_accumulation_over_pipeline_schedules(fused):
    """Two schedules per optimizer step (Trainer(grad_accum_steps=2) passes loss_scale=1/2 into the pipelined forward)."""
    torch.manual_seed(0)
    m = BloomForCausalLM(BloomConfig(**CFG)); state = copy.deepcopy(m.state_dict())
    data = [{"input_ids": torch.randint(0, 96, (4, 8))} for _ in range(4)]
    opt = FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5)
    for s in range(2):
        opt.zero_grad()
        for a in range(2):
            ids = data[2 * s + a]["input_ids"]
            (m(ids, labels=ids).loss / 2).backward()
        opt.step()
    spawn(run_pp_accumulation, world_size=2, state=state, data=data, ref_state={k: v.detach().clone() for k, v in m.state_dict().items()}, fused=fused)


def run_pp_eval(rank, world_size, port, state, ids, ref_loss, ref_eval):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 1)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    m = TensorParallel(m, ctx).parallelize()
    m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
    with torch.no_grad():
        l = m(ids, labels=ids).loss
    assert torch.allclose(l, ref_loss, atol=1e-5), (l, ref_loss)
    assert all(p.grad is None for p in m.parameters())
    t = Trainer(m, [{"input_ids": ids}], eval_loader=[{"input_ids": ids}, {"input_ids": ids.flip(0)}], optim=FusedAdam(m.parameters(), lr=1e-2), parallel_context=ctx)
    assert abs(t.evaluate() - ref_eval) < 1e-5
    ctx.destroy()
def test_pipelined_loss_under_no_grad_and_trainer_evaluate():
    torch.manual_seed(0)
    m = BloomForCausalLM(BloomConfig(**CFG)); ids = torch.randint(0, 96, (4, 8))
    with torch.no_grad():
        a = m(ids, labels=ids).loss; b = m(ids.flip(0), labels=ids.flip(0)).loss
    spawn(run_pp_eval, world_size=4, state=copy.deepcopy(m.state_dict()), ids=ids, ref_loss=a, ref_eval=((a + b) / 2).item())


def run_pp_dropped_step(rank, world_size, port, state, batches, ref_state, fused):
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    names = {id(p): n for n, p in m.named_parameters()}
    m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
    o = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5), ctx)
    # step 1 is abandoned after backward (what a loss scaler does on overflow): zero_grad BEFORE the next forward drops it
    loss = m(batches[0], labels=batches[0]).loss
    o.zero_grad()          # canonical position, between forward and backward: must NOT lose the schedule's gradients
    loss.backward()
    o.zero_grad()          # after backward: an ordinary zero_grad — the gradients are gone
    loss = m(batches[1], labels=batches[1]).loss
    loss.backward()
    o.step()
    for p in m._pg_pipeline_stage.parameters():
        assert torch.allclose(p.detach(), ref_state[names[id(p)]], atol=3e-5), names[id(p)]
    ctx.destroy()
@pytest.mark.parametrize("fused", [True, False])
def test_zero_grad_is_only_ignored_between_pipelined_forward_and_backward(fused):
    torch.manual_seed(0)
    m = BloomForCausalLM(BloomConfig(**CFG)); state = copy.deepcopy(m.state_dict())
    batches = [torch.randint(0, 96, (4, 8)) for _ in range(2)]
    opt = FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5)
    opt.zero_grad()
    m(batches[1], labels=batches[1]).loss.backward()     # only the second batch counts
    opt.step()
    spawn(run_pp_dropped_step, world_size=2, state=state, batches=batches,
          ref_state={k: v.detach().clone() for k, v in m.state_dict().items()}, fused=fused)
