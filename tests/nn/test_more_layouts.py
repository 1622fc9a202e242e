"""Layouts and corner cases found worth pinning while probing (CPU / gloo): stock optimizers behind generic ZeRO-1 at
dp=3, the fused ZeRO-1 path at dp=3, parameters that never receive a gradient, 🤗 Bloom (class-swap TP) with the fused
optimizer, GPT-2 with a vocabulary that needs padding through TP x PP training, deparallelize and cached generation."""
import copy

import pytest
import torch
from torch import nn

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel
from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)


# ------------------------------------------------------------------ ZeRO-1 at dp=3 with several optimizers
def _make_optim(name, params):
    if name == "sgd":
        return torch.optim.SGD(params, lr=0.1, momentum=0.9, nesterov=True)
    if name == "adamw":
        return torch.optim.AdamW(params, lr=1e-2, weight_decay=0.1)
    if name == "rmsprop":
        return torch.optim.RMSprop(params, lr=1e-3)
    if name == "fused":
        return FusedAdam(params, lr=1e-2)
    return FusedAdam(params, lr=1e-2, weight_decay=0.1, adamw=True)


def run_zero_dp3(rank, world_size, port, name, state, ids, ref_state):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, 3)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(_make_optim(name, model.parameters()), ctx)
    local = ids.chunk(3)[ctx.get_local_rank(ParallelMode.DATA)]
    for _ in range(3):
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
    for k, v in model.state_dict().items():
        assert torch.allclose(v, ref_state[k], atol=2e-5), k
    ctx.destroy()


@pytest.mark.parametrize("name", ["sgd", "adamw", "rmsprop", "fused", "fused_adamw"])
def test_zero1_with_three_replicas(name):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (6, 8))
    optim = _make_optim(name, model.parameters())
    for _ in range(3):
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
    spawn(run_zero_dp3, world_size=3, name=name, state=state, ids=ids,
          ref_state={k: v.clone() for k, v in model.state_dict().items()})


# ------------------------------------------------------------------ parameters without gradients
class _NetWithUnusedLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.unused, self.b = nn.Linear(8, 8), nn.Linear(8, 8), nn.Linear(8, 4)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def run_unused(rank, world_size, port, state, x, ref_state, fused):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    model = _NetWithUnusedLayer()
    model.load_state_dict(state)
    model = DataParallel(model, ctx).parallelize()
    inner = FusedAdam(model.parameters(), lr=1e-2) if fused else torch.optim.Adam(model.parameters(), lr=1e-2)
    optim = DistributedOptimizer(inner, ctx)
    for _ in range(3):
        loss = model(x.chunk(2)[rank]).pow(2).mean()
        optim.zero_grad()
        loss.backward()
        optim.step()
    for k, v in model.state_dict().items():
        assert torch.allclose(v, ref_state[k], atol=1e-5), k
    ctx.destroy()


@pytest.mark.parametrize("fused", [True, False])
def test_parameters_that_never_get_a_gradient_do_not_stall_the_reducer(fused):
    torch.manual_seed(0)
    model = _NetWithUnusedLayer()
    state = copy.deepcopy(model.state_dict())
    x = torch.randn(8, 8)
    optim = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(3):
        loss = model(x).pow(2).mean()
        optim.zero_grad()
        loss.backward()
        optim.step()
    spawn(run_unused, world_size=2, state=state, x=x, ref_state={k: v.clone() for k, v in model.state_dict().items()},
          fused=fused)


# ------------------------------------------------------------------ 🤗 Bloom, class-swap TP x DP, fused ZeRO-1
def _hf_bloom():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    return HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))


def run_hf_fused(rank, world_size, port, state, ids, ref_losses):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 2)
    model = _hf_bloom()
    model.load_state_dict(state)
    model.train()
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    got = []
    for _ in ref_losses:
        loss = model(input_ids=local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        got.append(loss.item())
    t = torch.tensor(got)
    torch.distributed.all_reduce(t)
    for a, b in zip((t / world_size).tolist(), ref_losses):
        assert abs(a - b) < 2e-3, (t, ref_losses)
    ctx.destroy()


def test_hf_bloom_with_the_fused_optimizer():
    torch.manual_seed(0)
    model = _hf_bloom()
    model.train()
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    optim, ref = torch.optim.Adam(model.parameters(), lr=1e-2), []
    for _ in range(3):
        optim.zero_grad()
        total = 0.0
        for chunk in ids.chunk(2):
            loss = model(input_ids=chunk, labels=chunk).loss / 2
            loss.backward()
            total += loss.item()
        optim.step()
        ref.append(total)
    spawn(run_hf_fused, world_size=4, state=state, ids=ids, ref_losses=ref)


# ------------------------------------------------------------------ GPT-2, padded vocabulary, TP x PP, export
GPT2 = dict(vocab_size=97, hidden_size=32, n_layer=4, n_head=4, n_positions=16)


def run_gpt2_round_trip(rank, world_size, port, state, ids, ref_losses, ref_generated):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 1)
    model = GPT2LMHeadModel(GPT2Config(**GPT2))
    model.load_state_dict(state)
    tp = TensorParallel(model, ctx)
    model = tp.parallelize()
    pp = PipelineParallel(model, num_microbatches=2, parallel_context=ctx)
    model = pp.parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    for want in ref_losses:
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        assert abs(loss.item() - want) < 2e-4, (loss.item(), ref_losses)
    model = pp.deparallelize()
    model = tp.deparallelize()
    assert torch.equal(model.generate(ids[:, :5], max_new_tokens=4), ref_generated)   # the trained, re-assembled model
    ctx.destroy()


def test_gpt2_padded_vocabulary_trains_under_tp_pp_and_exports():
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(**GPT2))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 97, (4, 8))
    optim, ref = FusedAdam(model.parameters(), lr=1e-2), []
    for _ in range(3):
        optim.zero_grad()
        total = 0.0
        for chunk in ids.chunk(2):
            loss = model(chunk, labels=chunk).loss / 2
            loss.backward()
            total += loss.item()
        optim.step()
        ref.append(total)
    spawn(run_gpt2_round_trip, world_size=4, state=state, ids=ids, ref_losses=ref,
          ref_generated=model.generate(ids[:, :5], max_new_tokens=4))


# ------------------------------------------------------------------ zero_grad() first, accumulation under no_sync()
def _hf_bloom_accum():
    from transformers import BloomConfig, BloomForCausalLM
    return BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
def run_zero_grad_first(rank, world_size, port, tp, fused, state, ids, ref_state):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 2)
    m = _hf_bloom_accum(); m.load_state_dict(state); m.train()
    names = {id(p): n for n, p in m.named_parameters()}
    m = TensorParallel(m, ctx).parallelize(); m = DataParallel(m, ctx).parallelize()
    opt = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5), ctx)
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    for step in range(2):
        mbs = local.chunk(2)
        opt.zero_grad()
        with m.no_sync():
            (m(input_ids=mbs[0], labels=mbs[0]).loss / 2).backward()
        (m(input_ids=mbs[1], labels=mbs[1]).loss / 2).backward()
        opt.step()
    for p in m.parameters():
        n = names.get(id(p))
        if n is not None and p.shape == ref_state[n].shape:
            assert torch.allclose(p.detach(), ref_state[n], atol=3e-5), n
    ctx.destroy()
@pytest.mark.parametrize("tp,fused", [(1, False), (1, True), (2, False), (2, True)])
def test_zero_grad_before_the_first_forward_and_no_sync_accumulation(tp, fused):
    """`optim.zero_grad()` first (the usual PyTorch order) must find and build the DataParallel reducer; micro-steps under
    `no_sync()` accumulate locally — 🤗 Bloom class-swap TP x DP, stock SGD and the fused ZeRO-1 optimizer."""
    torch.manual_seed(0)
    m = _hf_bloom_accum(); m.train(); state = copy.deepcopy(m.state_dict())
    ids = torch.randint(0, 96, (8, 8))
    opt = FusedAdam(m.parameters(), lr=1e-2, eps=1e-3) if fused else torch.optim.SGD(m.parameters(), lr=0.5)
    for step in range(2):
        opt.zero_grad()
        for rep in ids.chunk(2):
            for mb in rep.chunk(2):
                (m(input_ids=mb, labels=mb).loss / 4).backward()
        opt.step()
    spawn(run_zero_grad_first, world_size=2 * tp, tp=tp, fused=fused, state=state, ids=ids, ref_state={k: v.detach().clone() for k, v in m.state_dict().items()})


# ------------------------------------------------------------------ 🤗 families through all three wrappers
def _build_family(family):
    import transformers as T
    if family == "bloom":
        return T.BloomForCausalLM(T.BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    if family == "gpt2":
        return T.GPT2LMHeadModel(T.GPT2Config(vocab_size=96, n_embd=32, n_layer=4, n_head=4, n_positions=32, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0))
    return T.LlamaForCausalLM(T.LlamaConfig(vocab_size=96, hidden_size=32, intermediate_size=64, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=32, tie_word_embeddings=True))
def run_hf_3d(rank, world_size, port, family, fused, state, ids, ref_losses):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 2)
    m = _build_family(family); m.load_state_dict(state); m.train()
    m = TensorParallel(m, ctx).parallelize()
    m = PipelineParallel(m, num_microbatches=2, parallel_context=ctx).parallelize()
    m = DataParallel(m, ctx).parallelize()
    opt = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.Adam(m.parameters(), lr=1e-2), ctx)
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    got = []
    for _ in ref_losses:
        loss = m(input_ids=local, labels=local).loss
        opt.zero_grad(); loss.backward(); opt.step(); got.append(loss.item())
    t = torch.tensor(got); torch.distributed.all_reduce(t); t /= world_size
    for a, b in zip(t.tolist(), ref_losses):
        assert abs(a - b) < 2e-3, (family, t.tolist(), ref_losses)
    ctx.destroy()
@pytest.mark.parametrize("family,fused", [("bloom", False), ("bloom", True), ("gpt2", False), ("llama", True)])
def test_hf_families_through_tp_pp_dp_training(family, fused):
    """🤗 Bloom / GPT-2 / LLaMA, class-swap TP x pipeline stages x DP, stock Adam (reads ``.grad``) and the fused ZeRO-1 optimizer."""
    torch.manual_seed(0)
    m = _build_family(family); m.train(); state = copy.deepcopy(m.state_dict())
    ids = torch.randint(0, 96, (8, 8))
    opt = torch.optim.Adam(m.parameters(), lr=1e-2); ref = []
    chunks = [mb for rep in ids.chunk(2) for mb in rep.chunk(2)]
    for _ in range(3):
        opt.zero_grad(); tot = 0
        for c in chunks:
            l = m(input_ids=c, labels=c).loss / len(chunks); l.backward(); tot += l.item()
        opt.step(); ref.append(tot)
    spawn(run_hf_3d, world_size=8, family=family, fused=fused, state=state, ids=ids, ref_losses=ref)


# ------------------------------------------------------------------ 🤗 Bloom + MoE, replicated tokens
from pipegoose_b200.nn import ExpertParallel  # noqa: E402
from pipegoose_b200.nn.expert_parallel import ExpertLoss, SwitchNoisePolicy, Top1Router  # noqa: E402


def _hf_bloom_for_moe():
    import transformers as T
    return T.BloomForCausalLM(T.BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
def run_hf_moe(rank, world_size, port, tp, dp, state, gate_state, ids, out_file):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    m = _hf_bloom_for_moe(); m.load_state_dict(state)
    router = Top1Router(SwitchNoisePolicy(), 4, 32); router.load_state_dict(gate_state)
    m = ExpertParallel(m, 4, mapping=[1], router=router, parallel_context=ctx).parallelize()
    layer = m.transformer.h[1].mlp
    first = ctx.get_local_rank(ParallelMode.TENSOR) * len(layer.experts)
    for i, e in enumerate(layer.experts):
        g = torch.Generator().manual_seed(500 + first + i)
        for p in e.parameters():
            p.data = p.data + 0.05 * torch.randn(p.shape, generator=g)
    m = TensorParallel(m, ctx).parallelize(); m = DataParallel(m, ctx).parallelize(); m.eval()
    opt = DistributedOptimizer(torch.optim.Adam(m.parameters(), lr=1e-2), ctx)
    loss_fn = ExpertLoss(lambda out: out.loss, aux_weight=0.01, z_weight=0.001)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    losses = []
    for _ in range(3):
        loss = loss_fn(m(input_ids=local, labels=local)); opt.zero_grad(); loss.backward(); opt.step(); losses.append(loss.item())
    t = torch.tensor(losses); torch.distributed.all_reduce(t)
    if rank == 0: torch.save(t / world_size, out_file)
    ctx.destroy()
def test_hf_bloom_moe_with_sharded_experts_trains_like_unsharded(tmp_path):
    """Replicated tokens (class-swap TP), distinct experts sharded over the tensor group, DP + generic ZeRO-1 Adam: the
    gradients that flow back into the tokens and the router through the local experts are completed over the group."""
    torch.manual_seed(0)
    state = copy.deepcopy(_hf_bloom_for_moe().state_dict()); gate_state = copy.deepcopy(Top1Router(SwitchNoisePolicy(), 4, 32).state_dict())
    ids = torch.randint(0, 96, (4, 8)); res = {}
    for name, (tp, dp) in {"a": (1, 2), "b": (2, 2)}.items():
        f = str(tmp_path / name); spawn(run_hf_moe, world_size=tp * dp, tp=tp, dp=dp, state=state, gate_state=gate_state, ids=ids, out_file=f); res[name] = torch.load(f)
    assert torch.allclose(res["a"], res["b"], atol=2e-4), res
    assert res["a"][-1] < res["a"][0]


# ------------------------------------------------------------------ fast-path MoE, tensor parallelism only
from pipegoose_b200.nn.expert_parallel import Top2Router  # noqa: E402


def run_moe_tp_only(rank, world_size, port, tp, stock, top2, state, gate_state, ids, out_file):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    R = Top2Router if top2 else Top1Router
    router = R(SwitchNoisePolicy(), 4, 32); router.load_state_dict(gate_state)
    m = ExpertParallel(m, 4, mapping=[0, 1], router=router, parallel_context=ctx).parallelize()
    for li in (0, 1):
        layer = m.transformer.h[li].mlp
        first = ctx.get_local_rank(ParallelMode.TENSOR) * len(layer.experts)
        for i, e in enumerate(layer.experts):
            g = torch.Generator().manual_seed(500 + 10 * li + first + i)
            for p in e.parameters():
                p.data = p.data + 0.05 * torch.randn(p.shape, generator=g)
    m = TensorParallel(m, ctx).parallelize(); m.eval()
    opt = DistributedOptimizer(torch.optim.Adam(m.parameters(), lr=1e-2) if stock else FusedAdam(m.parameters(), lr=1e-2), ctx)
    loss_fn = ExpertLoss(lambda out: out.loss, aux_weight=0.01, z_weight=0.001)
    losses = []
    for _ in range(3):
        loss = loss_fn(m(ids, labels=ids)); opt.zero_grad(); loss.backward(); opt.step(); losses.append(loss.item())
    if rank == 0: torch.save(torch.tensor(losses), out_file)
    ctx.destroy()
@pytest.mark.parametrize("stock,top2", [(False, False), (True, False), (False, True)])
def test_fast_path_moe_without_data_parallel_is_tp_invariant(tmp_path, stock, top2):
    """tp = 1, 2, 4 (one expert per rank at tp=4: some ranks' experts get no token in a step — the backward collectives
    must still line up), no DataParallel reducer (the router's autograd gradients are folded before the tensor-group sum),
    fused and stock optimizer, Top-1 and Top-2: identical loss trajectories."""
    torch.manual_seed(0)
    R = Top2Router if top2 else Top1Router
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict()); gate_state = copy.deepcopy(R(SwitchNoisePolicy(), 4, 32).state_dict())
    ids = torch.randint(0, 96, (4, 8)); res = {}
    for tp in (1, 2, 4):
        f = str(tmp_path / f"t{tp}"); spawn(run_moe_tp_only, world_size=tp, tp=tp, stock=stock, top2=top2, state=state, gate_state=gate_state, ids=ids, out_file=f); res[tp] = torch.load(f)
    assert torch.allclose(res[1], res[2], atol=2e-4) and torch.allclose(res[1], res[4], atol=2e-4), res


# ------------------------------------------------------------------ experts that idle on one replica only
from pipegoose_b200.nn import ExpertParallel as _EP  # noqa: E402,F401


class _RouteByReplica(torch.nn.Module):
    """Routes every token of DP replica r to expert r: each replica leaves the other experts without gradients."""
    def __init__(self, e): super().__init__(); self.e = e
    def forward(self, x): return torch.full((x.reshape(-1, x.shape[-1]).shape[0],), self.e, dtype=torch.long)
def run_replica_local_experts(rank, world_size, port, fused, state, ids, ref_state):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, 2)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    m = ExpertParallel(m, 2, mapping=[0, 1], router=_RouteByReplica(rank), parallel_context=ctx).parallelize()
    m = DataParallel(m, ctx, bucket_size_mb=0.01).parallelize()
    opt = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.Adam(m.parameters(), lr=1e-2), ctx)
    local = ids.chunk(2)[rank]
    for _ in range(3):
        loss = m(local, labels=local).loss; opt.zero_grad(); loss.backward(); opt.step()
    assert len(m._pg_grad_reducer.buckets) > 3
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    other = flat.clone(); torch.distributed.all_reduce(other)
    assert torch.allclose(other / 2, flat, atol=1e-6), "replicas diverged"
    for k, v in m.state_dict().items():
        assert torch.allclose(v, ref_state[k], atol=3e-5), k
    ctx.destroy()
@pytest.mark.parametrize("fused", [True, False])
def test_parameters_unused_on_one_replica_only_keep_the_collective_order(fused):
    """Every replica routes its tokens to a different expert: each leaves other experts without gradients, so buckets
    complete in different orders on different replicas — the reducer must still issue its collectives in one order."""
    torch.manual_seed(0)
    base = BloomForCausalLM(BloomConfig(**CFG)); state = copy.deepcopy(base.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    # single-process reference: same MoE model, replica r's half of the batch goes to expert r
    from pipegoose_b200.testing.utils import find_free_port  # noqa: E402
    ctx = init_parallel_context(0, 1, find_free_port(), 1, 1, 1)
    class Split(torch.nn.Module):
        def forward(self, x):
            n = x.reshape(-1, x.shape[-1]).shape[0]; return (torch.arange(n) >= n // 2).long()
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    m = ExpertParallel(m, 2, mapping=[0, 1], router=Split(), parallel_context=ctx).parallelize()
    opt = FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(3):
        opt.zero_grad()
        loss = m(ids, labels=ids).loss; loss.backward(); opt.step()
    ref_state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ctx.destroy()
    spawn(run_replica_local_experts, world_size=2, fused=fused, state=state, ids=ids, ref_state=ref_state)


# ------------------------------------------------------------------ TP x DP MoE with idle experts
from pipegoose_b200.testing.utils import find_free_port  # noqa: E402


class _ConstRouter(torch.nn.Module):
    def __init__(self, e): super().__init__(); self.e = e
    def forward(self, x): return torch.full((x.reshape(-1, x.shape[-1]).shape[0],), self.e, dtype=torch.long)
class _SplitRouter(torch.nn.Module):
    def __init__(self, a, b): super().__init__(); self.a, self.b = a, b
    def forward(self, x):
        n = x.reshape(-1, x.shape[-1]).shape[0]; return torch.where(torch.arange(n) >= n // 2, self.b, self.a)
def _distinct_experts(m, ctx):
    for li in (0, 1):
        layer = m.transformer.h[li].mlp
        first = ctx.get_local_rank(ParallelMode.TENSOR) * len(layer.experts)
        for i, e in enumerate(layer.experts):
            g = torch.Generator().manual_seed(900 + 10 * li + first + i)
            for p in e.parameters(): p.data = p.data + 0.05 * torch.randn(p.shape, generator=g)
def run_idle_experts(rank, world_size, port, experts_of_replica, fused, state, ids, ref_state):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 2)
    dp_rank = ctx.get_local_rank(ParallelMode.DATA)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    names = {id(p): n for n, p in m.named_parameters()}
    m = ExpertParallel(m, 4, mapping=[0, 1], router=_ConstRouter(experts_of_replica[dp_rank]), parallel_context=ctx).parallelize()
    _distinct_experts(m, ctx)
    m = TensorParallel(m, ctx).parallelize(); m = DataParallel(m, ctx, bucket_size_mb=0.01).parallelize()
    opt = DistributedOptimizer(FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.Adam(m.parameters(), lr=1e-2), ctx)
    local = ids.chunk(2)[dp_rank]
    for _ in range(3):
        loss = m(local, labels=local).loss; opt.zero_grad(); loss.backward(); opt.step()
    for p in m.parameters():
        n = names.get(id(p))
        if n is not None and n in ref_state and p.shape == ref_state[n].shape:
            assert torch.allclose(p.detach(), ref_state[n], atol=3e-5), n
    ctx.destroy()
@pytest.mark.parametrize("experts_of_replica,fused", [((0, 1), True), ((0, 3), True), ((2, 1), False)])
def test_tp_dp_moe_with_experts_that_idle_on_some_ranks(experts_of_replica, fused):
    """Each replica sends all its tokens to ONE expert: some tensor ranks own no active expert (their backward must still
    run the exchange collectives) and experts idle on one replica only (bucket order) — against the single-process model."""
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    ids = torch.randint(0, 96, (4, 8))
    ctx = init_parallel_context(0, 1, find_free_port(), 1, 1, 1)
    m = BloomForCausalLM(BloomConfig(**CFG)); m.load_state_dict(state)
    m = ExpertParallel(m, 4, mapping=[0, 1], router=_SplitRouter(*experts_of_replica), parallel_context=ctx).parallelize()
    _distinct_experts(m, ctx)
    opt = FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(3):
        loss = m(ids, labels=ids).loss; opt.zero_grad(); loss.backward(); opt.step()
    ref_state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ctx.destroy()
    spawn(run_idle_experts, world_size=4, experts_of_replica=experts_of_replica, fused=fused, state=state, ids=ids, ref_state=ref_state)
