"""Elastic resume: a ZeRO-1 (FusedAdam) optimizer checkpoint written by ``dp`` replicas is loaded by a job with another
number of replicas (nn/utils.py::load_training_state -> reshard_fused_state) and training continues exactly — the global
batch is the same, only its split over replicas changes."""
import os

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.nn.utils import from_pretrained, load_training_state, reshard_fused_state, save_pretrained, save_training_state
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

GLOBAL_BATCH = 8


def _build(ctx):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model = TensorParallel(model, ctx).parallelize()
    if ctx.pipeline_parallel_size > 1:
        from pipegoose_b200.nn import PipelineParallel

        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    return model, DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2, weight_decay=0.01, adamw=True), ctx)


def _step(model, optim, ctx, i):
    ids = torch.randint(0, 96, (GLOBAL_BATCH, 8), generator=torch.Generator().manual_seed(100 + i))
    dp, r = ctx.get_world_size(ParallelMode.DATA), ctx.get_local_rank(ParallelMode.DATA)
    mine = ids.chunk(dp)[r]                                   # the same global batch whatever the replica count
    loss = model(mine, labels=mine).loss
    optim.zero_grad()
    loss.backward()
    optim.step()
    total = loss.detach().clone()
    dist.all_reduce(total, group=ctx.get_group(ParallelMode.DATA))
    return float(total) / dp


def _full_params(model, ctx):
    """{name: full tensor} of a TP-sharded model, gathered by hand (tp <= 2 here: rank order = tp rank)."""
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def run_write(rank, world_size, port, tp, dp, ckp, out, pp=1):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model, optim = _build(ctx)
    for i in range(2):
        _step(model, optim, ctx, i)
    save_pretrained(model, ckp_path=ckp, parallel_context=ctx)
    save_training_state(optim, ckp_path=ckp, parallel_context=ctx, step=2, extra={"tokens": 128})
    losses = [_step(model, optim, ctx, i) for i in (2, 3)]    # the uninterrupted run goes on
    if ctx.get_local_rank(ParallelMode.DATA) == 0:
        torch.save({"losses": losses, "params": _full_params(model, ctx)},
                   os.path.join(out, f"want_tp{ctx.get_local_rank(ParallelMode.TENSOR)}_pp{ctx.get_local_rank(ParallelMode.PIPELINE)}.pt"))
    ctx.destroy()


def run_resume(rank, world_size, port, tp, dp, ckp, out, warm, pp=1):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model, optim = _build(ctx)
    if warm:
        _step(model, optim, ctx, 99)          # optimizer state already laid out (and dirty) before the load
    from_pretrained(model, ckp_path=ckp, parallel_context=ctx)
    meta = load_training_state(optim, ckp_path=ckp, parallel_context=ctx)
    assert meta["step"] == 2 and meta["extra"] == {"tokens": 128} and meta["resharded_from_dp"] == 2
    losses = [_step(model, optim, ctx, i) for i in (2, 3)]
    want = torch.load(os.path.join(out, f"want_tp{ctx.get_local_rank(ParallelMode.TENSOR)}_pp{ctx.get_local_rank(ParallelMode.PIPELINE)}.pt"))
    assert losses == pytest.approx(want["losses"], abs=2e-5), (losses, want["losses"])
    for k, v in _full_params(model, ctx).items():
        assert torch.allclose(v, want["params"][k], atol=2e-5), k
    assert optim.optim._step == 4
    ctx.destroy()


@pytest.mark.parametrize("tp,new_dp,warm", [(1, 1, False), (1, 4, False), (2, 1, True), (1, 4, True)])
def test_zero1_checkpoint_of_two_replicas_resumes_on_another_replica_count(tmp_path, tp, new_dp, warm):
    ckp, out = str(tmp_path / "ckpt"), str(tmp_path)
    spawn(run_write, world_size=tp * 2, tp=tp, dp=2, ckp=ckp, out=out)
    spawn(run_resume, world_size=tp * new_dp, tp=tp, dp=new_dp, ckp=ckp, out=out, warm=warm)


def test_pipeline_stages_resume_on_another_replica_count(tmp_path):
    """PP2 x DP2 -> PP2 x DP1: every stage re-cuts the optimizer state of ITS parameters (the other stage's zero-size
    stand-ins are in the optimizer's parameter list but not in the flat state)."""
    ckp, out = str(tmp_path / "ckpt"), str(tmp_path)
    spawn(run_write, world_size=4, tp=1, dp=2, ckp=ckp, out=out, pp=2)
    spawn(run_resume, world_size=2, tp=1, dp=1, ckp=ckp, out=out, warm=False, pp=2)


def test_reshard_fused_state_unit():
    """Two old replicas own halves of two buckets of a 3-parameter buffer; the new layout moves the parameters (other
    padding) and has one owner."""
    old_index = [(0, 4), (4, 2), (8, 4)]                       # flat: p0[0:4] p1[4:6] pad[6:8] p2[8:12]
    full = torch.arange(12, dtype=torch.float32)

    def shard(segs, scale):
        return {"step": 7, "segments": segs, "param_groups": [{"lr": 0.5}],
                "master": torch.cat([full[s:e] for s, e in segs]) * scale,
                "exp_avg": torch.cat([full[s:e] for s, e in segs]) * 10 * scale,
                "exp_avg_sq": torch.cat([full[s:e] for s, e in segs]) * 100 * scale}

    olds = [shard([(0, 4), (8, 10)], 1.0), shard([(4, 8), (10, 12)], 1.0)]
    new_index = [(0, 4), (16, 2), (32, 4)]
    got = reshard_fused_state(olds, old_index, [(0, 48)], new_index, 48)
    assert got["step"] == 7 and got["segments"] == [(0, 48)] and got["param_groups"] == [{"lr": 0.5}]
    assert got["master"][0:4].tolist() == [0, 1, 2, 3] and got["master"][16:18].tolist() == [4, 5]
    assert got["master"][32:36].tolist() == [8, 9, 10, 11] and got["master"][4:16].abs().sum() == 0
    assert got["exp_avg_sq"][33] == 900
    two = reshard_fused_state(olds, old_index, [(0, 2), (32, 34)], new_index, 48)
    assert two["exp_avg"].tolist() == [0, 10, 80, 90]
    with pytest.raises(ValueError, match="do not cover"):
        reshard_fused_state(olds[:1], old_index, [(0, 48)], new_index, 48)
    with pytest.raises(ValueError, match="elements in the checkpoint"):
        reshard_fused_state(olds, old_index, [(0, 48)], [(0, 4), (16, 3), (32, 4)], 48)
    # parameters neither job holds on this rank (another pipeline stage's) are skipped; held by one job only: refused
    skip = reshard_fused_state(olds, old_index + [None], [(0, 48)], new_index + [None], 48)
    assert torch.equal(skip["master"], got["master"])
    with pytest.raises(ValueError, match="one of the two jobs only"):
        reshard_fused_state(olds, old_index + [None], [(0, 48)], new_index + [(40, 2)], 48)


def run_refused(rank, world_size, port, ckp):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)       # another TENSOR size: not re-cuttable
    model, optim = _build(ctx)
    with pytest.raises(ValueError):
        load_training_state(optim, ckp_path=ckp, parallel_context=ctx)
    ctx.destroy()


def test_other_layout_changes_are_still_refused(tmp_path):
    ckp = str(tmp_path / "ckpt")
    spawn(run_write, world_size=2, tp=1, dp=2, ckp=ckp, out=str(tmp_path))
    spawn(run_refused, world_size=2, ckp=ckp)


def run_trainer_phase(rank, world_size, port, dp, path, ckp, out, phase):
    """phase "full": 6 steps uninterrupted at this dp; "first": 3 steps + checkpoint; "rest": resume and finish."""
    from pipegoose_b200.trainer import Trainer
    from pipegoose_b200.utils.data import TokenFileDataset, build_dataloader

    ctx = init_parallel_context(rank, world_size, port, 1, 1, dp)
    model, optim = _build(ctx)
    loader = build_dataloader(TokenFileDataset(path, seq_len=8), ctx, batch_size=GLOBAL_BATCH // dp, shuffle=True)
    kw = {}
    if phase == "first":
        kw = dict(checkpoint_dir=ckp, checkpoint_every=3, max_steps=3)
    elif phase == "rest":
        kw = dict(checkpoint_dir=ckp, resume=True)
    state = Trainer(model, loader, optim=optim, parallel_context=ctx, num_epochs=2, **kw).fit()
    assert state.step == (3 if phase == "first" else 6)
    if rank == 0 and phase != "first":
        torch.save({k: v.detach().clone() for k, v in model.state_dict().items()}, os.path.join(out, f"{phase}.pt"))
    ctx.destroy()


def test_trainer_resumes_on_another_replica_count_with_the_same_global_batches(tmp_path):
    """Trainer + build_dataloader: 3 steps on 2 replicas (4 sequences each), checkpoint, 3 more steps on ONE replica (8
    sequences): the sharded sampler hands out the same global batches in both layouts, so the run ends where the
    uninterrupted 2-replica run ends."""
    from pipegoose_b200.utils.data import write_token_file

    path, ckp, out = str(tmp_path / "tokens.bin"), str(tmp_path / "ckpt"), str(tmp_path)
    write_token_file(path, torch.randint(0, 96, (24 * 8,), generator=torch.Generator().manual_seed(3)))   # 24 sequences: 3 steps/epoch
    spawn(run_trainer_phase, world_size=2, dp=2, path=path, ckp=ckp, out=out, phase="full")
    spawn(run_trainer_phase, world_size=2, dp=2, path=path, ckp=ckp, out=out, phase="first")
    spawn(run_trainer_phase, world_size=1, dp=1, path=path, ckp=ckp, out=out, phase="rest")
    want, got = torch.load(os.path.join(out, "full.pt")), torch.load(os.path.join(out, "rest.pt"))
    for k, v in want.items():
        assert torch.allclose(got[k], v, atol=2e-5), k
