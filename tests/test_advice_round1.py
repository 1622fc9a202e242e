"""Regression tests for the round-1 advisor findings."""
import copy
import random

import pytest
import torch

from pipegoose_b200.distributed import functional as F
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)


# ------------------------------------------------------------------ bare torch.optim under DataParallel
def run_dp_bare(rank, world_size, port, state, ids, ref):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = DataParallel(model, ctx).parallelize()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)   # NOT wrapped in DistributedOptimizer
    local = ids.chunk(world_size)[rank]
    for _ in range(3):
        loss = model(local, labels=local).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
    got = model.state_dict()
    for k, v in ref.items():
        assert torch.allclose(got[k], v, atol=1e-5), k
    ctx.destroy()


def test_data_parallel_with_a_bare_torch_optimizer_does_not_accumulate_across_steps():
    torch.manual_seed(0)
    m = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(m.state_dict())
    ids = torch.randint(0, 96, (4, 16))
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    for _ in range(3):
        loss = m(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
    spawn(run_dp_bare, world_size=2, state=state, ids=ids, ref=m.state_dict())


# ------------------------------------------------------------------ two synced backward() calls in one window
def run_two_backwards(rank, world_size, port, tp, dp, state, ids, ref_grads):
    """G1 + G2 accumulated by two SYNCED backward passes (no no_sync) must equal the single-process sum: the
    tensor-group SUM of the sequence-parallel partial gradients may not be applied to G1 twice."""
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    if dp > 1:
        model = DataParallel(model, ctx).parallelize()
    optim = FusedAdam(model.parameters(), lr=1e-3)   # consumes main_grad: no .grad materialisation
    r = ctx.get_local_rank(ParallelMode.DATA)
    a, b = ids[0].chunk(dp)[r], ids[1].chunk(dp)[r]
    optim.zero_grad()
    model(a, labels=a).loss.backward()
    model(b, labels=b).loss.backward()
    checked = 0
    for name, p in model.named_parameters():
        if getattr(p, "tp_partial_grad", False):
            assert torch.allclose(p.main_grad, ref_grads[name], atol=2e-5), name
            checked += 1
    assert checked > 0
    ctx.destroy()


@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2)])
def test_second_synced_backward_sums_only_its_own_partial_gradients(tp, dp):
    torch.manual_seed(1)
    m = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(m.state_dict())
    ids = torch.randint(0, 96, (2, 4, 16))
    m(ids[0], labels=ids[0]).loss.backward()
    m(ids[1], labels=ids[1]).loss.backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters()}
    spawn(run_two_backwards, world_size=tp * dp, tp=tp, dp=dp, state=state, ids=ids, ref_grads=ref)


def run_zero_double_sync(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    model = DataParallel(BloomForCausalLM(BloomConfig(**CFG)), ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-3), ctx)
    ids = torch.randint(0, 96, (2, 16))
    loss = model(ids, labels=ids).loss
    optim.zero_grad()
    loss.backward()
    with pytest.raises(RuntimeError, match="no_sync"):
        model(ids, labels=ids).loss.backward()   # a second reduce-scatter of the same window would be wrong: refuse
    ctx.destroy()


def test_zero1_refuses_a_second_synced_backward():
    spawn(run_zero_double_sync, world_size=2)


# ------------------------------------------------------------------ functional async collectives
def run_async_collectives(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    x = torch.full((2, 3), float(rank + 1))
    out, work = F.all_gather(x, dim=1, async_op=True, parallel_context=ctx, parallel_mode=ParallelMode.TENSOR)
    work.wait()
    want = torch.cat([torch.full((2, 3), float(r + 1)) for r in range(world_size)], dim=1)
    assert torch.equal(out, want)
    src = torch.arange(4.0) + rank
    keep = src.clone()
    res, work = F.reduce_scatter(src, dim=0, async_op=True, parallel_context=ctx, parallel_mode=ParallelMode.TENSOR)
    work.wait()
    assert torch.equal(src, keep), "reduce_scatter must not modify its input"
    total = sum(torch.arange(4.0) + r for r in range(world_size))
    n = 4 // world_size
    assert torch.equal(res, total[rank * n:(rank + 1) * n])
    ctx.destroy()


def test_async_all_gather_along_inner_dim_and_gloo_reduce_scatter():
    spawn(run_async_collectives, world_size=2)


# ------------------------------------------------------------------ strict checkpoint loading
def run_strict_load(rank, world_size, port, path):
    from pipegoose_b200.nn.utils import from_pretrained, save_pretrained

    ctx = init_parallel_context(rank, world_size, port, 1, 1, 1)
    small = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=1, n_head=4))
    save_pretrained(small, ckp_path=path, parallel_context=ctx)
    big = BloomForCausalLM(BloomConfig(**CFG))   # one more block than the checkpoint
    with pytest.raises(KeyError, match="lacks"):
        from_pretrained(big, ckp_path=path, parallel_context=ctx)
    from_pretrained(big, ckp_path=path, parallel_context=ctx, strict=False)
    wide = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=64, n_layer=1, n_head=4))
    with pytest.raises(ValueError, match="shape mismatch"):
        from_pretrained(wide, ckp_path=path, parallel_context=ctx)
    ctx.destroy()


def test_from_pretrained_is_strict(tmp_path):
    spawn(run_strict_load, world_size=1, path=str(tmp_path))


# ------------------------------------------------------------------ trainer
class _Loader:
    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)


def test_evaluate_restores_training_mode_and_stage():
    from pipegoose_b200.trainer import Trainer
    from pipegoose_b200.trainer.state import TrainerStage

    model = BloomForCausalLM(BloomConfig(**CFG))
    batches = [{"input_ids": torch.randint(0, 96, (2, 8))} for _ in range(2)]
    tr = Trainer(model, _Loader(batches), eval_loader=_Loader(batches), optim=torch.optim.SGD(model.parameters(), lr=0.1))
    model.train()
    tr.state.stage = TrainerStage.TRAINING
    tr.evaluate()
    assert model.training and tr.state.stage == TrainerStage.TRAINING
    model.eval()
    tr.evaluate()
    assert not model.training


def run_shuffled_resume(rank, world_size, port, path):
    """A run cut after 3 of 6 steps and resumed sees the same shuffled batches as an uninterrupted run."""
    from torch.utils.data import DataLoader

    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 1, 1)
    data = [{"input_ids": torch.full((8,), i, dtype=torch.long)} for i in range(12)]

    def make(seed):
        torch.manual_seed(seed)
        random.seed(seed)
        model = BloomForCausalLM(BloomConfig(**CFG))
        optim = FusedAdam(model.parameters(), lr=1e-3)
        loader = DataLoader(data, batch_size=2, shuffle=True)
        return model, optim, loader

    def run(max_steps, resume, seen, seed=7):
        model, optim, loader = make(seed)
        orig = model.forward

        def spy_forward(*a, **k):
            seen.append(k["input_ids"][:, 0].tolist())
            return orig(*a, **k)

        model.forward = spy_forward
        tr = Trainer(model, loader, optim=optim, parallel_context=ctx, checkpoint_dir=path, checkpoint_every=3,
                     resume=resume, max_steps=max_steps)
        tr.fit()
        return model

    full = []
    ref = run(6, False, full)
    import shutil

    shutil.rmtree(path)
    first, second = [], []
    run(3, False, first)
    got = run(6, True, second, seed=12345)   # a fresh process does not share the generator state of the interrupted one
    assert first + second == full, (first, second, full)
    for (k, a), (_, b) in zip(ref.state_dict().items(), got.state_dict().items()):
        assert torch.allclose(a, b, atol=1e-6), k
    ctx.destroy()


def test_resume_with_a_shuffled_dataloader_replays_the_same_batches(tmp_path):
    spawn(run_shuffled_resume, world_size=1, path=str(tmp_path / "ckpt"))


def run_two_backwards_stock_optimizer(rank, world_size, port, state, ids, ref_params):
    """Sequence-parallel model + plain torch.optim (no flat main grads), two synced backward passes, one step."""
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt.zero_grad()
    model(ids[0], labels=ids[0]).loss.backward()
    model(ids[1], labels=ids[1]).loss.backward()
    opt.step()
    for name, p in model.named_parameters():
        if getattr(p, "tp_partial_grad", False):     # replicated parameters: compare in full
            assert torch.allclose(p.detach(), ref_params[name], atol=2e-5), name
    ctx.destroy()


def test_stock_optimizer_on_the_sequence_parallel_model_two_backwards_per_step():
    torch.manual_seed(2)
    m = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(m.state_dict())
    ids = torch.randint(0, 96, (2, 4, 16))
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    opt.zero_grad()
    m(ids[0], labels=ids[0]).loss.backward()
    m(ids[1], labels=ids[1]).loss.backward()
    opt.step()
    spawn(run_two_backwards_stock_optimizer, world_size=2, state=state, ids=ids,
          ref_params={n: p.detach().clone() for n, p in m.named_parameters()})
