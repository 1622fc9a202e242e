"""More compositions (CPU / gloo): checkpoints of pipeline stages, FusedAdam's AdamW / weight-decay modes against
torch, 🤗 Bloom under TensorParallel (class-swap) x PipelineParallel, experts with tensor parallelism enabled."""
import copy
import os

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import ExpertParallel, PipelineParallel, TensorParallel
from pipegoose_b200.nn.expert_parallel import SwitchNoisePolicy, Top1Router
from pipegoose_b200.nn.utils import from_pretrained, save_pretrained
from pipegoose_b200.optim import FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=4, n_head=4)


def run_pp_checkpoint(rank, world_size, port, state, ckp_path):
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    stage = model._pg_pipeline_stage
    want = {k: v.clone() for k, v in stage.state_dict().items()}
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    assert os.path.exists(os.path.join(ckp_path, f"pytorch_model_tp_0_pp_{rank}.bin"))
    with torch.no_grad():
        for p in stage.parameters():
            p.zero_()
    from_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    for k, v in stage.state_dict().items():
        assert torch.equal(v, want[k]), k
    ctx.destroy()


def test_checkpoint_of_pipeline_stages(tmp_path):
    torch.manual_seed(0)
    spawn(run_pp_checkpoint, world_size=2, state=BloomForCausalLM(BloomConfig(**CFG)).state_dict(),
          ckp_path=str(tmp_path / "ckpt"))


@pytest.mark.parametrize("adamw,wd", [(False, 0.0), (False, 0.1), (True, 0.1)])
def test_fused_adam_modes_match_torch(adamw, wd):
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    mine = copy.deepcopy(ref)
    topt = (torch.optim.AdamW if adamw else torch.optim.Adam)(ref.parameters(), lr=1e-2, weight_decay=wd)
    fopt = FusedAdam(mine.parameters(), lr=1e-2, weight_decay=wd, adamw=adamw)
    x = torch.randn(16, 8)
    for _ in range(4):
        for model, opt in ((ref, topt), (mine, fopt)):
            opt.zero_grad()
            model(x).pow(2).mean().backward()
            opt.step()
    for a, b in zip(ref.parameters(), mine.parameters()):
        assert torch.allclose(a, b, atol=1e-5)


@pytest.mark.parametrize("sharded", [False, True])
def test_fused_adam_parameter_groups_match_torch(sharded):
    """Per-group hyper-parameters (no weight decay and a different lr for biases) — also when this rank only owns a
    slice of the flat buffer that cuts through parameters (ZeRO-1 segments)."""
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    mine = copy.deepcopy(ref)

    def groups(model):
        decay = [p for p in model.parameters() if p.dim() >= 2]
        rest = [p for p in model.parameters() if p.dim() < 2]
        return [{"params": decay, "weight_decay": 0.1}, {"params": rest, "weight_decay": 0.0, "lr": 3e-2}]

    topt = torch.optim.AdamW(groups(ref), lr=1e-2)
    fopt = FusedAdam(groups(mine), lr=1e-2, adamw=True)
    assert all(p._pg_fused_optim for p in mine.parameters())
    if sharded:
        n = fopt.ensure_flat().numel
        fopt.set_shard(n // 4, n // 4 + n // 2)
    before = [p.detach().clone() for p in mine.parameters()]
    x = torch.randn(16, 8)
    for _ in range(1 if sharded else 4):   # (a lone shard owner sees stale peers' slices after the first step)
        for model, opt in ((ref, topt), (mine, fopt)):
            opt.zero_grad()
            model(x).pow(2).mean().backward()
            opt.step()
    assert len({gi for _, _, _, gi in fopt._plan()}) == 2
    flat = fopt.flat
    s, e = fopt._segments[0]
    for a, b, b0 in zip(ref.parameters(), mine.parameters(), before):
        o, cnt = flat.param_range(b)
        idx = torch.arange(o, o + cnt)
        owned = ((idx >= s) & (idx < e)).view_as(b)
        assert torch.allclose(a[owned], b[owned], atol=1e-5)        # owned elements follow torch's AdamW
        assert torch.equal(b[~owned], b0[~owned])                   # the rest is another rank's to update


def _hf_bloom():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    return HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))


def run_hf_tp_pp(rank, world_size, port, state, ids, ref_loss):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 1)
    model = _hf_bloom()
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    out = model(ids, labels=ids)
    if ctx.is_last_rank(ParallelMode.PIPELINE):
        assert torch.allclose(out.loss, ref_loss, atol=1e-5), (out.loss, ref_loss)
    out.loss.backward()
    grads = [p.grad for p in model._pg_pipeline_stage.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    ctx.destroy()


def test_hf_bloom_tensor_x_pipeline_parallel_loss():
    torch.manual_seed(0)
    model = _hf_bloom()
    ids = torch.randint(0, 96, (4, 8))
    loss = torch.stack([model(input_ids=c, labels=c).loss for c in ids.chunk(2)]).mean()
    spawn(run_hf_tp_pp, world_size=4, state=copy.deepcopy(model.state_dict()), ids=ids, ref_loss=loss.detach())


def run_experts_with_tp(rank, world_size, port, state, gate_state, ids, ref_loss):
    """``enable_tensor_parallelism=True``: every rank holds ALL experts and the experts' linears are tensor-parallel
    (reference nn/expert_parallel/layers.py:28-31) — the loss equals the single-process MoE model's."""
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    model = _hf_bloom()
    model.load_state_dict(state)
    router = Top1Router(SwitchNoisePolicy(), 2, 32)
    router.load_state_dict(gate_state)
    model = ExpertParallel(model, 2, mapping=[0], router=router, enable_tensor_parallelism=True,
                           parallel_context=ctx).parallelize()
    layer = model.transformer.h[0].mlp
    assert len(layer.experts.experts if hasattr(layer.experts, "experts") else layer.experts) == 2
    model = TensorParallel(model, ctx).parallelize()
    model.eval()
    with torch.no_grad():
        loss = model(input_ids=ids, labels=ids).loss
    assert torch.allclose(loss, ref_loss, atol=1e-5), (loss, ref_loss)
    ctx.destroy()


def test_expert_layer_with_tensor_parallel_experts(tmp_path):
    torch.manual_seed(0)
    model = _hf_bloom()
    state = copy.deepcopy(model.state_dict())
    gate_state = copy.deepcopy(Top1Router(SwitchNoisePolicy(), 2, 32).state_dict())
    ids = torch.randint(0, 96, (2, 8))
    f = str(tmp_path / "ref.pt")

    def ref_run(rank, world_size, port):
        ctx = init_parallel_context(rank, world_size, port, 1, 1, 1)
        m = _hf_bloom()
        m.load_state_dict(state)
        r = Top1Router(SwitchNoisePolicy(), 2, 32)
        r.load_state_dict(gate_state)
        m = ExpertParallel(m, 2, mapping=[0], router=r, enable_tensor_parallelism=True, parallel_context=ctx).parallelize()
        m.eval()
        with torch.no_grad():
            torch.save(m(input_ids=ids, labels=ids).loss, f)
        ctx.destroy()

    # the single-process MoE reference runs in-process through a world-size-1 context
    port = 29000 + os.getpid() % 2000
    ref_run(0, 1, port)
    spawn(run_experts_with_tp, world_size=2, state=state, gate_state=gate_state, ids=ids, ref_loss=torch.load(f))


def run_moe_pp(rank, world_size, port, state, gate_state, ids, ref_loss):
    """Switch-MoE blocks inside pipeline stages (ExpertParallel -> PipelineParallel): the scheduled step computes the
    same language-model loss as the unpartitioned MoE model."""
    ctx = init_parallel_context(rank, world_size, port, 1, 2, 1)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    router = Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"])
    router.load_state_dict(gate_state)
    model = ExpertParallel(model, 2, mapping=[0, 3], router=router, parallel_context=ctx).parallelize()
    model.eval()
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    out = model(ids, labels=ids)
    if ctx.is_last_rank(ParallelMode.PIPELINE):
        assert torch.allclose(out.loss, ref_loss, atol=1e-5), (out.loss, ref_loss)
    out.loss.backward()
    grads = [p.grad for p in model._pg_pipeline_stage.parameters() if p.requires_grad]
    assert any(g is not None and g.abs().sum() > 0 for g in grads)
    assert all(g is None or torch.isfinite(g).all() for g in grads)
    ctx.destroy()


def test_moe_blocks_inside_pipeline_stages(tmp_path):
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    gate_state = copy.deepcopy(Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"]).state_dict())
    ids = torch.randint(0, 96, (4, 8))
    f = str(tmp_path / "ref.pt")

    def ref_run(port):
        ctx = init_parallel_context(0, 1, port, 1, 1, 1)
        m = BloomForCausalLM(BloomConfig(**CFG))
        m.load_state_dict(state)
        r = Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"])
        r.load_state_dict(gate_state)
        m = ExpertParallel(m, 2, mapping=[0, 3], router=r, parallel_context=ctx).parallelize()
        m.eval()
        with torch.no_grad():
            torch.save(torch.stack([m(c, labels=c).loss for c in ids.chunk(2)]).mean(), f)
        ctx.destroy()

    ref_run(31000 + os.getpid() % 2000)
    spawn(run_moe_pp, world_size=2, state=state, gate_state=gate_state, ids=ids, ref_loss=torch.load(f))


def run_moe_4d(rank, world_size, port, tp, pp, dp, state, gate_state, ids, ref_losses, tol):
    """ExpertParallel -> TensorParallel -> PipelineParallel -> DataParallel -> ZeRO-1: every stage back-propagates the
    router losses of its own MoE layers, so the trajectory follows the unpartitioned MoE model trained on
    LM + 0.01 aux + 0.001 z."""
    from pipegoose_b200.nn import DataParallel
    from pipegoose_b200.nn.expert_parallel import ExpertContext
    from pipegoose_b200.optim import DistributedOptimizer

    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    router = Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"])
    router.load_state_dict(gate_state)
    model = ExpertParallel(model, 2, mapping=[0, 3], router=router, parallel_context=ctx).parallelize()
    model.eval()
    model = TensorParallel(model, ctx).parallelize()
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    store = ExpertContext.get_instance()
    for want in ref_losses:
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        assert not store.aux_loss and not store.z_loss          # drained by the engine, nothing piles up
        t = loss.detach().float().reshape(1).clone()
        torch.distributed.all_reduce(t)
        assert abs(t.item() / world_size - want) < tol, (t.item() / world_size, ref_losses)
    ctx.destroy()


@pytest.mark.parametrize("tp,tol", [(1, 3e-4), (2, 3e-4)])
def test_moe_tensor_pipeline_data_parallel_training(tmp_path, tp, tol):
    # the pipeline reproduces the unpartitioned objective (router losses of every stage included); with tp=2 the
    # sequence-parallel MoE layers exchange tokens (all-gather / reduce-scatter) so routing statistics stay global
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    gate_state = copy.deepcopy(Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"]).state_dict())
    ids = torch.randint(0, 96, (8, 8))
    dp, n_mb = 2, 2
    f = str(tmp_path / "ref.pt")

    def ref_run(port):
        from pipegoose_b200.nn.expert_parallel import ExpertContext

        ctx = init_parallel_context(0, 1, port, 1, 1, 1)
        m = BloomForCausalLM(BloomConfig(**CFG))
        m.load_state_dict(state)
        r = Top1Router(SwitchNoisePolicy(), 2, CFG["hidden_size"])
        r.load_state_dict(gate_state)
        m = ExpertParallel(m, 2, mapping=[0, 3], router=r, parallel_context=ctx).parallelize()
        m.eval()
        opt = FusedAdam(m.parameters(), lr=1e-2)
        store = ExpertContext.get_instance()
        chunks = [mb for rep in ids.chunk(dp) for mb in rep.chunk(n_mb)]
        out = []
        for _ in range(3):
            opt.zero_grad()
            lm_total = 0.0
            for mb in chunks:
                lm = m(mb, labels=mb).loss
                total = lm + 0.01 * sum(store.pop_all_aux_loss()) + 0.001 * sum(store.pop_all_z_loss())
                (total / len(chunks)).backward()
                lm_total += lm.item() / len(chunks)
            opt.step()
            out.append(lm_total)
        torch.save(out, f)
        ctx.destroy()

    ref_run(33000 + os.getpid() % 2000)
    ref_losses = torch.load(f)
    assert ref_losses[-1] < ref_losses[0]
    spawn(run_moe_4d, world_size=tp * 2 * dp, tp=tp, pp=2, dp=dp, state=state, gate_state=gate_state, ids=ids,
          ref_losses=ref_losses, tol=tol)
