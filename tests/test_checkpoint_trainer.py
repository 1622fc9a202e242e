import copy
import os

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.nn.utils import from_pretrained, save_pretrained
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def run_checkpoint(rank, world_size, port, tp, dp, ckp_path):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    want = {k: v.clone() for k, v in model.state_dict().items()}
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    name = f"pytorch_model_tp_{ctx.get_local_rank(ParallelMode.TENSOR)}_pp_0.bin"
    assert os.path.exists(os.path.join(ckp_path, name))
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    from_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    for k, v in model.state_dict().items():
        assert torch.equal(v, want[k]), k
    ctx.destroy()


@pytest.mark.parametrize("world,tp,dp", [(1, 1, 1), (4, 2, 2)])
def test_save_and_load_sharded_checkpoint(tmp_path, world, tp, dp):
    spawn(run_checkpoint, world_size=world, tp=tp, dp=dp, ckp_path=str(tmp_path / "ckpt"))


def test_trainer_fits_and_logs():
    import io

    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.trainer import Callback, DistributedLogger, Trainer, TrainerStatus

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4))
    data = [{"input_ids": torch.randint(0, 64, (2, 8))} for _ in range(6)]
    events = []

    class Rec(Callback):
        def on_fit_start(self, trainer):
            events.append("start")

        def on_step_end(self, trainer, loss):
            events.append(float(loss))

        def on_fit_end(self, trainer):
            events.append("end")

    import json
    import tempfile

    from pipegoose_b200.trainer import JsonlLogger

    stream = io.StringIO()
    metrics_file = os.path.join(tempfile.mkdtemp(), "run", "metrics.jsonl")
    trainer = Trainer(model, data, optim=FusedAdam(model.parameters(), lr=1e-2), num_epochs=2, callbacks=[Rec()],
                      loggers=[DistributedLogger(stream=stream), JsonlLogger(metrics_file)], log_every=3, max_grad_norm=10.0)
    state = trainer.fit()
    rows = [json.loads(line) for line in open(metrics_file)]
    assert [r["step"] for r in rows] == [3, 6, 9, 12] and rows[-1]["loss"] < rows[0]["loss"]
    assert all({"tokens_per_s", "tokens_seen", "grad_norm", "lr"} <= set(r) for r in rows) and rows[0]["lr"] == 1e-2
    assert state.status is TrainerStatus.FINISHED and state.step == 12 and state.tokens_seen == 12 * 16
    assert events[0] == "start" and events[-1] == "end" and events[-2] < events[1]  # loss went down
    assert "tokens/s" in stream.getvalue()


def run_resume(rank, world_size, port, tp, dp, ckp_path):
    from pipegoose_b200.nn.utils import load_training_state, save_training_state
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam

    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    cfg = BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)
    ids = torch.randint(0, 96, (4, 8), generator=torch.Generator().manual_seed(7 + ctx.get_local_rank(ParallelMode.DATA)))

    def build():
        torch.manual_seed(0)
        model = BloomForCausalLM(cfg)
        model = TensorParallel(model, ctx).parallelize()
        model = DataParallel(model, ctx).parallelize()
        return model, DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)

    def step(model, optim):
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        return loss.item()

    model, optim = build()
    for _ in range(2):
        step(model, optim)
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    save_training_state(optim, ckp_path=ckp_path, parallel_context=ctx, step=2, extra={"tokens": 64})
    want = step(model, optim)  # the third step of the uninterrupted run

    model2, optim2 = build()
    step(model2, optim2)  # build the lazily created optimizer state, then overwrite everything
    from_pretrained(model2, ckp_path=ckp_path, parallel_context=ctx)
    meta = load_training_state(optim2, ckp_path=ckp_path, parallel_context=ctx)
    assert meta == {"step": 2, "extra": {"tokens": 64}}
    got = step(model2, optim2)
    assert abs(got - want) < 1e-5, (got, want)
    ctx.destroy()


@pytest.mark.parametrize("world,tp,dp", [(1, 1, 1), (4, 2, 2)])
def test_resume_from_sharded_optimizer_checkpoint(tmp_path, world, tp, dp):
    spawn(run_resume, world_size=world, tp=tp, dp=dp, ckp_path=str(tmp_path / "ckpt"))


def run_trainer_resume(rank, world_size, port, tp, dp, ckp_dir):
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    cfg = BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)
    g = torch.Generator().manual_seed(11 + ctx.get_local_rank(ParallelMode.DATA))
    data = [{"input_ids": torch.randint(0, 96, (2, 8), generator=g)} for _ in range(12)]  # 12 micro-batches = 6 steps

    def build(**kw):
        torch.manual_seed(0)
        model = BloomForCausalLM(cfg)
        model = TensorParallel(model, ctx).parallelize()
        model = DataParallel(model, ctx).parallelize()
        optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2, eps=1e-3), ctx)
        sched = torch.optim.lr_scheduler.LambdaLR(optim.optim, lambda step: 1.0 / (1 + step))
        trainer = Trainer(model, data, optim=optim, parallel_context=ctx, grad_accum_steps=2, max_grad_norm=0.05,
                          lr_scheduler=sched, **kw)
        return model, optim, trainer

    # uninterrupted run: 6 optimizer steps
    model_a, optim_a, trainer_a = build()
    state = trainer_a.fit()
    assert state.step == 6 and state.tokens_seen == 12 * 16 and state.last_grad_norm > 0.05
    assert abs(optim_a.optim.param_groups[0]["lr"] - 1e-2 / 7) < 1e-9
    # interrupted after 4 steps (checkpoint every 2) ...
    model_b, optim_b, trainer_b = build(checkpoint_dir=ckp_dir, checkpoint_every=2)
    trainer_b.train_loader = data[:8]
    assert trainer_b.fit().step == 4
    # ... and resumed by a fresh process-local model / optimizer / trainer
    model_c, optim_c, trainer_c = build(checkpoint_dir=ckp_dir, checkpoint_every=0, resume=True)
    state = trainer_c.fit()
    assert state.step == 6 and state.tokens_seen == 12 * 16
    for (n, a), (_, c) in zip(model_a.named_parameters(), model_c.named_parameters()):
        assert torch.allclose(a, c, atol=1e-6), n
    assert abs(optim_c.optim.param_groups[0]["lr"] - 1e-2 / 7) < 1e-9
    ctx.destroy()


def test_trainer_accumulates_clips_schedules_checkpoints_and_resumes(tmp_path):
    spawn(run_trainer_resume, world_size=4, tp=2, dp=2, ckp_dir=str(tmp_path / "run"))


def run_trainer_vs_reference(rank, world_size, port, tp, dp, state, data_by_dp, ref_state):
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2, eps=1e-3), ctx)
    sched = torch.optim.lr_scheduler.LambdaLR(optim.optim, lambda step: 1.0 / (1 + step))
    data = data_by_dp[ctx.get_local_rank(ParallelMode.DATA)]
    Trainer(model, data, optim=optim, parallel_context=ctx, grad_accum_steps=2, max_grad_norm=0.05, lr_scheduler=sched).fit()
    for p in model.parameters():
        n = names.get(id(p))
        if n is not None and p.shape == ref_state[n].shape:
            assert torch.allclose(p.detach(), ref_state[n], atol=3e-5), n
    ctx.destroy()


def test_trainer_matches_a_hand_written_single_process_loop():
    """Accumulation (no_sync) + clipping + LR schedule under TP x DP x ZeRO-1 against an independent reference loop."""
    from pipegoose_b200.optim import FusedAdam, clip_grad_norm_

    class One:
        def get_world_size(self, mode):
            return 1

    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    state = copy.deepcopy(model.state_dict())
    dp, accum, steps = 2, 2, 3
    g = torch.Generator().manual_seed(5)
    data_by_dp = [[{"input_ids": torch.randint(0, 96, (2, 8), generator=g)} for _ in range(accum * steps)] for _ in range(dp)]
    opt = FusedAdam(model.parameters(), lr=1e-2, eps=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 1.0 / (1 + step))
    for s in range(steps):
        opt.zero_grad()
        for a in range(accum):
            for r in range(dp):
                ids = data_by_dp[r][s * accum + a]["input_ids"]
                (model(ids, labels=ids).loss / (accum * dp)).backward()
        clip_grad_norm_(opt, 0.05, One())
        opt.step()
        sched.step()
    ref_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    spawn(run_trainer_vs_reference, world_size=4, tp=2, dp=dp, state=state, data_by_dp=data_by_dp, ref_state=ref_state)


def run_ckpt_layout(rank, world_size, port, ckp_dir):
    from pipegoose_b200.nn import DataParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    torch.manual_seed(0)
    model = DataParallel(BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4)), ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    data = [{"input_ids": torch.randint(0, 64, (2, 8))} for _ in range(6)]
    Trainer(model, data, optim=optim, parallel_context=ctx, checkpoint_dir=ckp_dir, checkpoint_every=2, keep_checkpoints=2).fit()
    torch.distributed.barrier()
    kept = sorted(d for d in os.listdir(ckp_dir) if d.startswith("step_"))
    assert kept == ["step_00000004", "step_00000006"], kept                     # the two newest complete checkpoints
    assert open(os.path.join(ckp_dir, "latest")).read() == "step_00000006"
    files = sorted(os.listdir(os.path.join(ckp_dir, "step_00000006")))
    assert files == ["optimizer_tp_0_pp_0_dp_0.bin", "optimizer_tp_0_pp_0_dp_1.bin", "pytorch_model_tp_0_pp_0.bin",
                         "pytorch_model_tp_0_pp_0.bin.layout.json"], files
    assert not [f for f in os.listdir(ckp_dir) if ".tmp." in f or f.startswith(".latest")]
    torch.distributed.barrier()   # every rank has looked at the directory before anyone adds to it
    # a half-written newer directory (crash before "latest" moved) is ignored by resume
    os.makedirs(os.path.join(ckp_dir, "step_00000008"), exist_ok=True)
    trainer = Trainer(model, data, optim=optim, parallel_context=ctx, checkpoint_dir=ckp_dir, resume=True)
    assert trainer.load_checkpoint() and trainer.state.step == 6
    ctx.destroy()


def test_trainer_checkpoint_directories_and_latest_marker(tmp_path):
    spawn(run_ckpt_layout, world_size=2, ckp_dir=str(tmp_path / "ckpt"))


def test_trainer_adds_and_drains_router_losses():
    from pipegoose_b200.nn import ExpertParallel
    from pipegoose_b200.nn.expert_parallel import ExpertContext, SwitchNoisePolicy, Top1Router
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.testing.utils import find_free_port
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(0, 1, find_free_port(), 1, 1, 1)
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=2, n_head=4))
    router = Top1Router(SwitchNoisePolicy(), 2, 32)
    model = ExpertParallel(model, 2, router=router, parallel_context=ctx).parallelize()
    data = [{"input_ids": torch.randint(0, 64, (2, 8))} for _ in range(4)]
    gate_before = router.gate.weight.detach().clone()
    Trainer(model, data, optim=FusedAdam(model.parameters(), lr=1e-2), parallel_context=ctx).fit()
    store = ExpertContext.get_instance()
    assert not store.aux_loss and not store.z_loss            # drained every step
    assert not torch.equal(router.gate.weight.detach(), gate_before)
    ctx.destroy()


def run_trainer_resume_pipeline(rank, world_size, port, ckp):
    from pipegoose_b200.nn import DataParallel, PipelineParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 2, 2)
    g = torch.Generator().manual_seed(3 + ctx.get_local_rank(ParallelMode.DATA))
    data = [{"input_ids": torch.randint(0, 96, (4, 8), generator=g)} for _ in range(6)]

    def build(**kw):
        torch.manual_seed(0)
        m = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
        m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
        m = DataParallel(m, ctx).parallelize()
        return m, Trainer(m, data, optim=DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2), ctx), parallel_context=ctx, **kw)

    ma, ta = build()
    ta.fit()
    mb, tb = build(checkpoint_dir=ckp, checkpoint_every=2, max_steps=4)
    tb.fit()
    mc, tc = build(checkpoint_dir=ckp, resume=True)
    assert tc.fit().step == 6
    for (n, a), (_, c) in zip(ma.named_parameters(), mc.named_parameters()):
        assert a.shape == c.shape and torch.allclose(a, c, atol=1e-6), n
    ctx.destroy()


def test_trainer_checkpoints_and_resumes_pipeline_stages(tmp_path):
    """PP2 x DP2 + ZeRO-1: the optimizer's parameter list holds the other stage's zero-size stand-ins, the flat state
    does not (the per-parameter index of the optimizer shard must cope)."""
    spawn(run_trainer_resume_pipeline, world_size=4, ckp=str(tmp_path / "run"))


def test_periodic_evaluation_leaves_training_untouched(tmp_path):
    """``eval_every``: evaluations at steps 2, 4 and at the end (5); the trained parameters equal those of a run without
    evaluations although the model uses dropout and the eval loader is a DataLoader (whose iterator draws from the RNG)."""
    import json

    from torch.utils.data import DataLoader

    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.trainer import Callback, JsonlLogger, Trainer

    cfg = BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4, hidden_dropout=0.1)
    g = torch.Generator().manual_seed(1)
    train = [{"input_ids": torch.randint(0, 64, (2, 8), generator=g)} for _ in range(5)]
    held_out = [{"input_ids": torch.randint(0, 64, (8,), generator=g)} for _ in range(6)]

    def run(**kw):
        torch.manual_seed(0)
        model = BloomForCausalLM(cfg)
        trainer = Trainer(model, train, optim=FusedAdam(model.parameters(), lr=1e-2), **kw)
        trainer.fit()
        return model, trainer

    plain, _ = run()
    seen = []

    class Rec(Callback):
        def on_evaluate(self, trainer, eval_loss):
            seen.append((trainer.state.step, eval_loss, trainer.module.training))

    path = str(tmp_path / "metrics.jsonl")
    evaluated, trainer = run(eval_loader=DataLoader(held_out, batch_size=3, shuffle=True), eval_every=2, callbacks=[Rec()],
                             loggers=[JsonlLogger(path)], log_every=100)
    assert [s for s, _, _ in seen] == [2, 4, 5] and all(training for _, _, training in seen)
    assert seen[-1][1] == trainer.state.last_eval_loss and seen[-1][1] < seen[0][1] + 0.5
    rows = [json.loads(line) for line in open(path)]
    assert [r["step"] for r in rows if "eval_loss" in r] == [2, 4, 5]
    for (n, a), (_, b) in zip(plain.named_parameters(), evaluated.named_parameters()):
        assert torch.equal(a, b), n


def run_dp_evaluate(rank, world_size, port, state, batches, want):
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=1, n_head=4))
    model.load_state_dict(state)
    model = DataParallel(TensorParallel(model, ctx).parallelize(), ctx).parallelize()
    mine = batches[rank::world_size]                               # replica 0 gets two batches, replica 1 one
    trainer = Trainer(model, mine, eval_loader=mine, optim=FusedAdam(model.parameters(), lr=1e-2), parallel_context=ctx)
    assert abs(trainer.evaluate() - want) < 1e-5                   # the mean over ALL replicas' batches, on every rank
    ctx.destroy()


def test_evaluate_averages_over_the_data_parallel_replicas():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=1, n_head=4))
    batches = [{"input_ids": torch.randint(0, 96, (2, 8))} for _ in range(3)]
    with torch.no_grad():
        want = sum(float(model(b["input_ids"], labels=b["input_ids"]).loss) for b in batches) / 3
    spawn(run_dp_evaluate, world_size=2, state=copy.deepcopy(model.state_dict()), batches=batches, want=want)


def run_skip_nonfinite(rank, world_size, port, fused):
    from pipegoose_b200.nn import DataParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    g = torch.Generator().manual_seed(4 + rank)
    batches = [{"input_ids": torch.randint(0, 96, (2, 8), generator=g)} for _ in range(4)]

    def run(data, poison_call):
        torch.manual_seed(0)
        model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=1, n_head=4))
        model = DataParallel(TensorParallel(model, ctx).parallelize(), ctx).parallelize()
        inner = FusedAdam(model.parameters(), lr=1e-2) if fused else torch.optim.Adam(model.parameters(), lr=1e-2)
        calls = [0]

        class Poisoned(Trainer):                # only replica 1's loss (hence gradient) is poisoned: every rank sees the NORM
            def _add_router_losses(self, loss):
                calls[0] += 1
                loss = super()._add_router_losses(loss)
                return loss * float("nan") if (calls[0] == poison_call and rank == 1) else loss

        trainer = Poisoned(model, data, optim=DistributedOptimizer(inner, ctx), parallel_context=ctx, max_grad_norm=1.0)
        state = trainer.fit()
        return model, state

    poisoned, state = run(batches, poison_call=2)
    assert state.step == 3 and state.skipped_steps == 1
    clean, state = run([batches[0], batches[2], batches[3]], poison_call=-1)
    assert state.step == 3 and state.skipped_steps == 0
    for (n, a), (_, b) in zip(poisoned.named_parameters(), clean.named_parameters()):
        assert torch.isfinite(a).all() and torch.allclose(a, b, atol=1e-6), n
    ctx.destroy()


@pytest.mark.parametrize("fused", [True, False])
def test_trainer_skips_a_step_with_a_non_finite_gradient_norm(fused):
    """One replica's gradient turns NaN in the second step: every rank drops that step (the norm is global), weights and
    Adam state stay as they were — the run ends where a run without that batch ends."""
    spawn(run_skip_nonfinite, world_size=2, fused=fused)


def test_fit_saves_the_last_steps_when_periodic_checkpoints_are_on(tmp_path):
    """5 steps, a checkpoint every 2: steps 2, 4 — and 5 when ``fit`` ends, so a finished run can be continued or exported
    from its real end state (nothing extra when the run ends on a periodic checkpoint, or when checkpointing is off)."""
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.testing.utils import find_free_port
    from pipegoose_b200.trainer import Trainer

    ctx = init_parallel_context(0, 1, find_free_port(), 1, 1, 1)
    try:
        def run(n_batches, ckp, every):
            torch.manual_seed(0)
            model = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=32, n_layer=1, n_head=4))
            data = [{"input_ids": torch.randint(0, 64, (2, 8))} for _ in range(n_batches)]
            Trainer(model, data, optim=FusedAdam(model.parameters(), lr=1e-2), parallel_context=ctx, checkpoint_dir=ckp,
                    checkpoint_every=every, keep_checkpoints=5).fit()
            return sorted(d for d in os.listdir(ckp) if d.startswith("step_")) if os.path.isdir(ckp) else []

        assert run(5, str(tmp_path / "a"), 2) == ["step_00000002", "step_00000004", "step_00000005"]
        assert open(str(tmp_path / "a" / "latest")).read() == "step_00000005"
        assert run(4, str(tmp_path / "b"), 2) == ["step_00000002", "step_00000004"]
        assert run(3, str(tmp_path / "c"), 0) == []
    finally:
        ctx.destroy()
