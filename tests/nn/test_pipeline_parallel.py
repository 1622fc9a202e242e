"""Pipeline parallelism: schedules, micro-batching, structural partitioning, and the p2p engine
(GPipe driven by autograd, and the scheduled 1F1B training step) against sequential execution."""
import copy

import pytest
import torch
from torch import nn

from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import PipelineParallel
from pipegoose_b200.nn.pipeline_parallel import microbatch
from pipegoose_b200.nn.pipeline_parallel._job.job_type import JobType
from pipegoose_b200.nn.pipeline_parallel._utils import get_partition_idx, is_last_stage
from pipegoose_b200.nn.pipeline_parallel.partitioner import UniformPartitioner
from pipegoose_b200.nn.pipeline_parallel.pipeline_context import PipelineContext, TrainingState
from pipegoose_b200.nn.pipeline_parallel.scheduler import GPipeScheduler, OneFOneBScheduler, SchedulerType, get_scheduler
from pipegoose_b200.testing.utils import init_parallel_context, init_pipeline_context, spawn


@pytest.mark.parametrize("m,n", [(4, 2), (5, 3), (2, 4), (8, 4)])
def test_gpipe_schedule(m, n):
    sch = GPipeScheduler(m, n)
    clocks = sch.get_schedules()
    assert sch.total_clock_cycles == 2 * (m + n - 1)
    assert sch.total_forward_clock_cycles == sch.total_backward_clock_cycles == m + n - 1
    fwd = sch.get_forward_schedules()
    for c, tasks in enumerate(fwd):
        for t in tasks:
            assert t.job_type is JobType.FORWARD and t.microbatch_idx + t.partition_idx == c
    seen = [(t.job_type, t.microbatch_idx, t.partition_idx) for clock in clocks for t in clock]
    assert len(seen) == len(set(seen)) == 2 * m * n
    assert get_scheduler(SchedulerType.GPIPE) is GPipeScheduler


@pytest.mark.parametrize("m,n", [(4, 2), (8, 4), (3, 4), (6, 3)])
def test_1f1b_schedule(m, n):
    sch = OneFOneBScheduler(m, n)
    for p in range(n):
        order = sch.get_stage_order(p)
        assert len(order) == 2 * m
        done_f, done_b, alive, peak = set(), set(), 0, 0
        for t in order:
            if t.job_type is JobType.FORWARD:
                assert t.microbatch_idx == len(done_f)
                done_f.add(t.microbatch_idx)
                alive += 1
            else:
                assert t.microbatch_idx in done_f and t.microbatch_idx == len(done_b)
                done_b.add(t.microbatch_idx)
                alive -= 1
            peak = max(peak, alive)
        assert peak <= min(n - p, m)  # bounded activation memory: the point of 1F1B
    clocks = sch.get_schedules()
    assert sum(len(c) for c in clocks) == 2 * m * n
    assert len(clocks) == 2 * (m + n - 1)  # same bubble as GPipe


def test_microbatch_split_counts_not_sizes():
    inputs = {"input_ids": torch.arange(36 * 4).view(36, 4), "attention_mask": torch.ones(36, 4)}
    for n in (6, 4, 5):
        mbs = microbatch.split(inputs, n)
        assert len(mbs) == n
        assert all(set(mb) == set(inputs) for mb in mbs)
        assert sum(mb["input_ids"].shape[0] for mb in mbs) == 36
    assert torch.equal(torch.cat([mb["input_ids"] for mb in microbatch.split(inputs, 5)]), inputs["input_ids"])


def run_partitioner(rank, world_size, port, pp, state, ids, ref_logits):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=6, n_head=4))
    model.load_state_dict(state)
    stages = UniformPartitioner(model, ctx).split(["input_ids"])
    assert len(stages) == pp
    assert sum(len(s.h) for s in stages) == 6 and all(len(s.h) >= 1 for s in stages)
    x = ids
    for s in stages:  # chaining the partitions reproduces the full model
        x = s(x, batch_seq=tuple(ids.shape))
    assert torch.allclose(x, ref_logits, atol=1e-5)
    assert get_partition_idx(ctx) == rank and is_last_stage(ctx) == (rank == pp - 1)
    ctx.destroy()


@pytest.mark.parametrize("pp", [2, 4])
def test_partitioner_reproduces_model(pp):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=6, n_head=4))
    ids = torch.randint(0, 96, (2, 8))
    with torch.no_grad():
        ref = model(ids).logits
    spawn(run_partitioner, world_size=pp, pp=pp, state=copy.deepcopy(model.state_dict()), ids=ids, ref_logits=ref)


def run_pipeline_context(rank, world_size, port):
    pctx, ctx = init_pipeline_context(rank, world_size, port, 1, world_size, 1, n_microbatches=4)
    assert pctx.partition_idx == rank and pctx.num_microbatches == 4
    assert type(pctx).get_context() is pctx
    assert pctx.is_first_stage == (rank == 0) and pctx.is_last_stage == (rank == world_size - 1)
    assert pctx.state is TrainingState.IDLE
    pctx.forward()
    assert pctx.state is TrainingState.FORWARD
    total = 0
    for clock, tasks in enumerate(pctx.get_schedule()):
        assert pctx.clock_idx == clock and all(t.partition_idx == rank for t in tasks)
        total += len(tasks)
    assert total == 8  # 4 forward + 4 backward tasks of this partition
    assert pctx.is_last_microbatch(3) and not pctx.is_last_microbatch(0)
    ctx.destroy()


def test_pipeline_context():
    spawn(run_pipeline_context, world_size=2)


def _sequential_model():
    torch.manual_seed(0)
    return nn.Sequential(*[nn.Sequential(nn.Linear(8, 8), nn.Tanh()) for _ in range(4)])


def run_gpipe_autograd(rank, world_size, port, pp, state, x, ref_outs, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = _sequential_model()
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=4, parallel_context=ctx, scheduler_type=SchedulerType.GPIPE).parallelize()
    outputs = model(x)  # reference usage: one output per micro-batch, user drives backward
    assert len(outputs) == 4
    if rank == pp - 1:
        for o, r in zip(outputs, ref_outs):
            assert torch.allclose(o, r, atol=1e-6)
    for o in outputs:
        o.sum().backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=1e-5), names[id(p)]
    ctx.destroy()


@pytest.mark.parametrize("pp", [2, 4])
def test_gpipe_forward_then_user_driven_backward(pp):
    model = _sequential_model()
    x = torch.randn(8, 8)
    outs = [model(c) for c in x.chunk(4)]
    for o in outs:
        o.sum().backward()
    spawn(run_gpipe_autograd, world_size=pp, pp=pp, state=copy.deepcopy(model.state_dict()), x=x,
          ref_outs=[o.detach() for o in outs], ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def run_1f1b_bloom(rank, world_size, port, pp, sched, state, ids, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=4, parallel_context=ctx, scheduler_type=sched).parallelize()
    out = model(ids, labels=ids)
    assert torch.allclose(out.loss, ref_loss, atol=1e-5)  # computed on the last stage, broadcast to every stage
    out.loss.backward()  # harmless: the schedule already ran backward
    for p in model._pg_pipeline_stage.parameters():
        n = names[id(p)]
        # the tied table lives on the first and the last stage: the engine sums both contributions
        assert torch.allclose(p.grad, ref_grads[n], atol=2e-5), n
    ctx.destroy()


@pytest.mark.parametrize("pp,sched", [(2, SchedulerType.ONE_F_ONE_B), (4, SchedulerType.ONE_F_ONE_B), (2, SchedulerType.GPIPE)])
def test_scheduled_training_step_matches_sequential(pp, sched):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (8, 8))
    # same loss definition as the pipeline: mean over micro-batches of the per-micro-batch mean loss
    losses = [model(c, labels=c).loss for c in ids.chunk(4)]
    loss = torch.stack(losses).mean()
    loss.backward()
    spawn(run_1f1b_bloom, world_size=pp, pp=pp, sched=sched, state=copy.deepcopy(model.state_dict()), ids=ids,
          ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def _hf_model(family):
    if family == "gpt2":
        from transformers import GPT2Config, GPT2LMHeadModel

        return GPT2LMHeadModel(GPT2Config(vocab_size=96, n_positions=32, n_embd=32, n_layer=6, n_head=4,
                                          resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0))
    if family == "llama":
        from transformers import LlamaConfig, LlamaForCausalLM

        return LlamaForCausalLM(LlamaConfig(vocab_size=96, hidden_size=32, intermediate_size=64, num_hidden_layers=6,
                                            num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=32,
                                            tie_word_embeddings=True))
    from transformers import BloomConfig as HFBloomConfig
    from transformers import BloomForCausalLM as HFBloom

    return HFBloom(HFBloomConfig(vocab_size=96, hidden_size=32, n_layer=6, n_head=4))


def run_hf_partitioner(rank, world_size, port, pp, family, state, ids, ref_logits, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = _hf_model(family)
    model.load_state_dict(state)
    stages = UniformPartitioner(model, ctx).split(["input_ids"])
    assert len(stages) == pp and sum(len(s.h if hasattr(s, "h") else s.layers) for s in stages) == 6
    x = ids
    with torch.no_grad():
        for s in stages:  # the reference's acceptance test: chained partitions reproduce the full model's logits
            x = s(x)
    assert torch.allclose(x, ref_logits, atol=1e-5)
    # and the pipeline engine trains it: scheduled 1F1B step == sequential micro-batched step
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    out = model(ids, labels=ids)
    if rank == pp - 1:
        assert torch.allclose(out.loss, ref_loss, atol=1e-5)
    out.loss.backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=2e-5), names[id(p)]
    ctx.destroy()


@pytest.mark.parametrize("family,pp", [("gpt2", 2), ("gpt2", 3), ("bloom", 2), ("bloom", 3), ("llama", 2), ("llama", 3)])
def test_partitioner_and_engine_on_hf_models(family, pp):
    torch.manual_seed(0)
    model = _hf_model(family).eval()
    ids = torch.randint(0, 96, (4, 8))
    with torch.no_grad():
        ref_logits = model(ids).logits
    losses = [model(input_ids=c, labels=c).loss for c in ids.chunk(2)]
    loss = torch.stack(losses).mean()
    loss.backward()
    spawn(run_hf_partitioner, world_size=pp, pp=pp, family=family, state=copy.deepcopy(model.state_dict()), ids=ids,
          ref_logits=ref_logits, ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def run_pp_deparallelize(rank, world_size, port, pp, state, ids, ref_logits):
    ctx = init_parallel_context(rank, world_size, port, 1, pp, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    wrapper = PipelineParallel(model, num_microbatches=2, parallel_context=ctx)
    model = wrapper.parallelize()
    assert sum(p.numel() == 0 for p in model.parameters()) > 0  # foreign stages were dropped
    model(ids, labels=ids)  # one scheduled step still works
    model = wrapper.deparallelize()
    for k, v in model.state_dict().items():
        assert torch.equal(v, state[k]), k
    assert model.lm_head.weight is model.transformer.word_embeddings.weight
    with torch.no_grad():
        assert torch.allclose(model(ids).logits, ref_logits, atol=1e-5)
    ctx.destroy()


@pytest.mark.parametrize("pp", [2, 4])
def test_pipeline_parallel_deparallelize_restores_the_model(pp):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (4, 8))
    with torch.no_grad():
        ref = model(ids).logits
    spawn(run_pp_deparallelize, world_size=pp, pp=pp, state=copy.deepcopy(model.state_dict()), ids=ids, ref_logits=ref)


def run_engine_facade(rank, world_size, port, state, ids, ref_loss):
    from pipegoose_b200.nn.pipeline_parallel.pipeline import _PipelineEngine
    from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType

    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    with pytest.raises(ValueError):
        _PipelineEngine(model, num_concurrent=2, max_concurrent=1, parallel_context=ctx)
    with pytest.raises(TypeError):
        _PipelineEngine(model, parallel_context=None)
    engine = _PipelineEngine(model, scheduler=SchedulerType.GPIPE, parallel_context=ctx, num_microbatches=2)
    assert engine.parallelize() is model
    loss = engine(ids, labels=ids).loss
    assert torch.allclose(loss.detach().float().cpu(), ref_loss, atol=1e-5)
    ctx.destroy()


def test_pipeline_engine_facade_runs_a_pipelined_step():
    """Reference nn/pipeline_parallel/pipeline.py (_PipelineEngine, a stub there) as a working front door."""
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (4, 8))
    ref = torch.stack([model(c, labels=c).loss for c in ids.chunk(2)]).mean().detach()
    spawn(run_engine_facade, world_size=2, state=copy.deepcopy(model.state_dict()), ids=ids, ref_loss=ref)


def run_uneven_microbatches(rank, world_size, port, state, ids, labels, n_mb, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 1, world_size, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = PipelineParallel(model, num_microbatches=n_mb, parallel_context=ctx).parallelize()
    out = model(ids, labels=labels)
    assert torch.allclose(out.loss, ref_loss, atol=1e-5), (out.loss, ref_loss)
    out.loss.backward()
    for p in model._pg_pipeline_stage.parameters():
        assert torch.allclose(p.grad, ref_grads[names[id(p)]], atol=2e-5), names[id(p)]
    ctx.destroy()


@pytest.mark.parametrize("batch,n_mb,pp", [(5, 2, 2), (6, 4, 2), (5, 2, 4), (7, 3, 4)])
def test_uneven_microbatches_and_ignored_labels_give_the_global_token_mean(batch, n_mb, pp):
    """Micro-batch losses are weighted by their share of target tokens: the pipelined loss and gradients equal the
    unpartitioned model's even when the batch does not split evenly and some labels are -100."""
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (batch, 8))
    labels = ids.clone()
    labels[0, 5:] = -100          # padding on one sequence
    loss = model(ids, labels=labels).loss
    loss.backward()
    spawn(run_uneven_microbatches, world_size=pp, state=copy.deepcopy(model.state_dict()), ids=ids, labels=labels, n_mb=n_mb,
          ref_loss=loss.detach(), ref_grads={n: p.grad.clone() for n, p in model.named_parameters()})


def test_partitioner_register_family_and_unknown_model_message():
    """A model family the partitioner does not know is refused with an explanation, and ``register_family`` teaches it."""
    import pytest
    from torch import nn

    from pipegoose_b200.nn.pipeline_parallel.partitioner import SequentialStage, UniformPartitioner

    class Odd(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder_blocks = nn.ModuleList([nn.Linear(8, 8) for _ in range(4)])

    class Ctx:
        pipeline_parallel_size = 2

    with pytest.raises(NotImplementedError, match="register_family"):
        UniformPartitioner(Odd(), Ctx()).split()
    saved = list(UniformPartitioner._FAMILIES)
    try:
        UniformPartitioner.register_family(
            lambda m: isinstance(m, Odd), lambda m: m.encoder_blocks,
            lambda m, a, b, first, last: SequentialStage(list(m.encoder_blocks[a:b])))
        stages = UniformPartitioner(Odd(), Ctx()).split()
        assert [len(s.layers) for s in stages] == [2, 2]
    finally:
        UniformPartitioner._FAMILIES[:] = saved


@pytest.mark.parametrize("family", ["bloom", "gpt2"])
def test_partitions_chain_in_the_reference_call_style(family):
    """``out = stage0(**inputs); out = stage1(*out); ...`` — how the reference's partitioner test drives the stages."""
    from transformers import BloomConfig as HFBloomConfig
    from transformers import BloomForCausalLM as HFBloom
    from transformers import GPT2Config, GPT2LMHeadModel

    class Ctx:
        pipeline_parallel_size = 3

    torch.manual_seed(0)
    if family == "bloom":
        model = HFBloom(HFBloomConfig(vocab_size=128, hidden_size=32, n_layer=6, n_head=4)).eval()
    else:
        model = GPT2LMHeadModel(GPT2Config(vocab_size=128, n_embd=32, n_layer=6, n_head=4)).eval()
    inputs = {"input_ids": torch.randint(0, 128, (2, 7)), "attention_mask": torch.ones(2, 7, dtype=torch.long)}
    want = model(**inputs).logits
    stages = UniformPartitioner(model, Ctx()).split()
    out = inputs
    for stage in stages:
        out = stage(*out) if type(out) in (list, tuple) else stage(**out)
    assert torch.allclose(out, want, atol=1e-6)
    # the library's own call style still gets plain tensors
    h = stages[0](inputs["input_ids"], attention_mask=inputs["attention_mask"])
    assert isinstance(h, torch.Tensor) and h.shape == (2, 7, 32)


@pytest.mark.parametrize("m,n", [(8, 2), (4, 4), (16, 4), (3, 2)])
def test_schedule_timing_model_bubble_and_live_activations(m, n):
    """``simulate`` / ``bubble_fraction`` / ``peak_live_microbatches``: the closed forms for equal stages."""
    for cls in (GPipeScheduler, OneFOneBScheduler):
        sch = cls(m, n)
        makespan, busy, start = sch.simulate(1.0, 2.0)
        assert makespan == pytest.approx(3.0 * (m + n - 1)) and busy == [3.0 * m] * n
        assert sch.bubble_fraction() == pytest.approx((n - 1) / (m + n - 1))
        assert len(start) == 2 * m * n
        # transfers stretch the critical path: 2 (n - 1) hops there and back
        assert sch.simulate(1.0, 2.0, transfer_cost=0.5)[0] >= makespan + 2 * (n - 1) * 0.5 - 1e-9
    assert [GPipeScheduler(m, n).peak_live_microbatches(p) for p in range(n)] == [m] * n
    assert [OneFOneBScheduler(m, n).peak_live_microbatches(p) for p in range(n)] == [min(n - p, m) for p in range(n)]


def run_forward_only_bloom(rank, world_size, port, tp, state, ids, ref_logits, ref_grads):
    """The reference's forward-only usage on the fused Bloom: ``outs = model(ids); for o in outs: o.sum().backward()``."""
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.nn import TensorParallel

    ctx = init_parallel_context(rank, world_size, port, tp, 2, 1)
    m = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    m.load_state_dict(state)
    names = {id(p): n for n, p in m.named_parameters()}
    m = TensorParallel(m, ctx).parallelize()
    m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
    outs = m(ids)
    if ctx.is_last_rank(ParallelMode.PIPELINE):
        got = torch.cat(list(outs), 0)
        assert got.shape == ref_logits.shape, got.shape          # (tokens gathered, vocabulary gathered)
        assert torch.allclose(got, ref_logits, atol=1e-5)
    for o in outs:
        o.sum().backward()
    if tp == 1:     # (sharded parameters: the shapes differ; the logits check above covers tp = 2)
        checked = 0
        for p in m._pg_pipeline_stage.parameters():
            g = p.grad if p.grad is not None else getattr(p, "main_grad", None)
            n = names[id(p)]
            if n in ("lm_head.weight", "transformer.word_embeddings.weight"):
                continue        # tied table: each stage holds its own contribution here (the engine sums them in training)
            assert g is not None and torch.allclose(g.to(ref_grads[n].dtype), ref_grads[n], atol=2e-5), n
            checked += 1
        assert checked > 0
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_forward_only_pipeline_returns_full_logits_and_backpropagates(tp):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    ids = torch.randint(0, 96, (4, 6))
    logits = model(ids).logits
    logits.sum().backward()
    grads = {n: (p.grad if p.grad is not None else p.main_grad).detach().clone() for n, p in model.named_parameters()}
    spawn(run_forward_only_bloom, world_size=2 * tp, tp=tp, state=copy.deepcopy(model.state_dict()), ids=ids,
          ref_logits=logits.detach(), ref_grads=grads)


def run_hf_generate_refused(rank, world_size, port):
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    from pipegoose_b200.nn import TensorParallel

    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    torch.manual_seed(0)
    model = HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model = TensorParallel(model, ctx, sequence_parallel=False).parallelize()
    wrapper = PipelineParallel(model, num_microbatches=2, parallel_context=ctx)
    model = wrapper.parallelize()
    ids = torch.randint(0, 96, (2, 5))
    with pytest.raises(NotImplementedError, match="not pipeline-aware"):
        model.generate(input_ids=ids, max_new_tokens=2)
    model = wrapper.deparallelize()                       # 🤗's own generate is back
    out = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=2, do_sample=False, pad_token_id=0)
    assert out.shape == (2, 7)
    ctx.destroy()


def test_generate_of_a_pipelined_hf_model_is_refused_with_directions():
    spawn(run_hf_generate_refused, world_size=2)
