"""🤗 Bloom -> fused path: ``TensorParallel(hf_bloom, ctx).parallelize()`` converts the model IN PLACE to the
pipegoose_b200 Bloom (same parameters, same names) and then takes the sequence-parallel path; losses, gradients,
one optimizer step and ``generate`` agree with the untouched 🤗 model."""
import copy

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomForCausalLM as FastBloom
from pipegoose_b200.models.bloom import convert_hf_bloom_, is_hf_bloom
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def _hf_bloom(**kw):
    from transformers import BloomConfig, BloomForCausalLM

    return BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4, **kw))


def test_in_place_conversion_keeps_parameters_and_matches_hf():
    torch.manual_seed(0)
    hf = _hf_bloom()
    ref = copy.deepcopy(hf)
    names = [n for n, _ in hf.named_parameters()]
    params = {n: p for n, p in hf.named_parameters()}
    assert is_hf_bloom(hf)
    fast = convert_hf_bloom_(hf)
    assert fast is hf and isinstance(hf, FastBloom) and not is_hf_bloom(hf)
    assert [n for n, _ in hf.named_parameters()] == names
    assert all(p is params[n] for n, p in hf.named_parameters())          # the very same Parameter objects
    assert set(hf.state_dict()) == set(ref.state_dict())
    ids = torch.randint(0, 96, (3, 10))
    out = hf(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids)
    want = ref(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids)
    assert torch.allclose(out.loss, want.loss, atol=1e-5)
    out.loss.backward()
    want.loss.backward()
    for (n, p), (_, q) in zip(hf.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5), n
    assert torch.allclose(hf(input_ids=ids).logits, ref(input_ids=ids).logits, atol=1e-4)
    assert torch.equal(hf.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=3),
                       ref.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=3, do_sample=False))


def test_models_the_fused_path_cannot_run_keep_the_class_swap_path():
    hf = _hf_bloom(apply_residual_connection_post_layernorm=True)
    with pytest.raises(ValueError, match="post_layernorm"):
        convert_hf_bloom_(hf)


def _partition_of(name, full, tp, r):
    if "query_key_value" in name or "dense_h_to_4h" in name:
        return full.chunk(tp, 0)[r]
    if name.endswith("self_attention.dense.weight") or name.endswith("dense_4h_to_h.weight"):
        return full.chunk(tp, 1)[r]
    if "word_embeddings.weight" in name or name == "lm_head.weight":
        return full.chunk(tp, 0)[r]
    return full


def run_hf_fast(rank, world_size, port, tp, dp, state, ids, ref_loss, ref_params, ref_tokens):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    model = _hf_bloom()
    model.load_state_dict(state)
    same = model
    model = TensorParallel(model, ctx, sequence_parallel=True).parallelize()   # what bf16 🤗 models get by default
    assert model is same and isinstance(model, FastBloom) and (model.tp is not None) == (tp > 1)
    gen_in = ids[:2, :6]
    assert torch.equal(model.generate(input_ids=gen_in, max_new_tokens=2), ref_tokens)
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(torch.optim.Adam(model.parameters(), lr=1e-3), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    out = model(input_ids=local, attention_mask=torch.ones_like(local), labels=local)
    optim.zero_grad()
    out.loss.backward()
    optim.step()
    import torch.distributed as dist

    t = out.loss.detach().clone()
    dist.all_reduce(t, group=ctx.get_group(ParallelMode.DATA))
    assert torch.allclose(t / dp, ref_loss, atol=1e-5)
    r = ctx.get_local_rank(ParallelMode.TENSOR)
    for name, p in model.named_parameters():
        want = _partition_of(name, ref_params[name], tp, r)
        assert p.shape == want.shape, name
        assert torch.allclose(p.detach(), want, atol=2e-5), name
    ctx.destroy()


@pytest.mark.parametrize("tp,dp", [(1, 1), (2, 1), (2, 2)])
def test_hf_bloom_takes_the_sequence_parallel_path(tp, dp):
    torch.manual_seed(0)
    model = _hf_bloom()
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    tokens = model.generate(input_ids=ids[:2, :6], attention_mask=torch.ones(2, 6, dtype=torch.long), max_new_tokens=2,
                            do_sample=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids).loss
    opt.zero_grad()
    loss.backward()
    opt.step()
    ref_params = {n: p.detach().clone() for n, p in model.named_parameters()}
    ref_params["lm_head.weight"] = ref_params["transformer.word_embeddings.weight"]
    spawn(run_hf_fast, world_size=tp * dp, tp=tp, dp=dp, state=state, ids=ids, ref_loss=loss.detach(),
          ref_params=ref_params, ref_tokens=tokens)


def test_default_picks_the_fast_path_for_bf16_hf_models_only():
    class Ctx:   # tp == 1: parallelize() only decides about the conversion
        tensor_parallel_size = 1

    fp32 = _hf_bloom()
    assert is_hf_bloom(TensorParallel(fp32, Ctx()).parallelize())               # reference-style handling
    bf16 = _hf_bloom().to(torch.bfloat16)
    assert isinstance(TensorParallel(bf16, Ctx()).parallelize(), FastBloom)     # ready for the kernels
    off = _hf_bloom().to(torch.bfloat16)
    assert is_hf_bloom(TensorParallel(off, Ctx(), sequence_parallel=False).parallelize())


def test_hidden_dropout_trains_through_the_composed_path():
    """``hidden_dropout > 0``: eval equals the 🤗 model, training applies dropout (stochastic, mean-preserving) and the
    gradients flow to every parameter."""
    torch.manual_seed(0)
    hf = _hf_bloom(hidden_dropout=0.2)
    ref = copy.deepcopy(hf)
    fast = convert_hf_bloom_(hf)
    ids = torch.randint(0, 96, (3, 10))
    fast.eval(), ref.eval()
    assert torch.allclose(fast(input_ids=ids, labels=ids).loss, ref(input_ids=ids, labels=ids).loss, atol=1e-5)
    fast.train()
    torch.manual_seed(1)
    a = fast(input_ids=ids, labels=ids).loss
    torch.manual_seed(2)
    b = fast(input_ids=ids, labels=ids).loss
    assert not torch.allclose(a, b)                      # different dropout masks
    a.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in fast.parameters())
    # with p -> 0 the composed path reproduces the fused one
    for blk in fast.transformer.h:
        blk.hidden_dropout = 1e-12
    fast.zero_grad()
    ref.train()
    assert torch.allclose(fast(input_ids=ids, labels=ids).loss, _no_dropout_loss(ref, ids), atol=1e-5)


def test_attention_dropout_trains_through_the_composed_path():
    """``attention_dropout > 0`` (refused until round 2): eval equals the 🤗 model, training drops attention probabilities
    (stochastic), gradients reach every parameter, p -> 0 reproduces the fused path."""
    torch.manual_seed(0)
    hf = _hf_bloom(attention_dropout=0.3)
    ref = copy.deepcopy(hf)
    fast = convert_hf_bloom_(hf)
    assert all(blk.attention_dropout == 0.3 for blk in fast.transformer.h)
    ids = torch.randint(0, 96, (3, 10))
    fast.eval(), ref.eval()
    assert torch.allclose(fast(input_ids=ids, labels=ids).loss, ref(input_ids=ids, labels=ids).loss, atol=1e-5)
    fast.train()
    torch.manual_seed(1)
    a = fast(input_ids=ids, labels=ids).loss
    torch.manual_seed(2)
    b = fast(input_ids=ids, labels=ids).loss
    assert not torch.allclose(a, b)
    a.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in fast.parameters())
    for blk in fast.transformer.h:
        blk.attention_dropout = 1e-12
    assert torch.allclose(fast(input_ids=ids, labels=ids).loss, _no_dropout_loss(ref, ids), atol=1e-5)


def test_attention_dropout_semantics():
    """The head-grouped path equals dropout applied to the full fp32 probabilities (same generator state), and the
    expectation over masks is the attention without dropout."""
    from pipegoose_b200.ops.attention import alibi_attention, alibi_attention_reference
    from pipegoose_b200.ops import kernels as K

    B, S, H, D = 2, 12, 4, 8
    torch.manual_seed(0)
    qkv = torch.randn(B * S, H * 3 * D)
    slopes = K.alibi_slopes(H)
    torch.manual_seed(5)
    got = alibi_attention(qkv, slopes, B, S, H, D, dropout_p=0.25)
    torch.manual_seed(5)
    want = alibi_attention_reference(qkv, slopes, B, S, H, D, dropout_p=0.25)
    assert torch.allclose(got, want, atol=1e-5)
    clean = alibi_attention_reference(qkv, slopes, B, S, H, D)
    assert not torch.allclose(got, clean, atol=1e-3)
    mean = torch.stack([alibi_attention(qkv, slopes, B, S, H, D, dropout_p=0.25) for _ in range(400)]).mean(0)
    assert (mean - clean).abs().max() < 0.25


def _no_dropout_loss(hf_model, ids):
    hf_model.eval()   # 🤗 dropout off == p -> 0
    return hf_model(input_ids=ids, labels=ids).loss


def run_tp_cached_generate(rank, world_size, port, tp, state, ids, ref_tokens):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    model = _hf_bloom()
    model.load_state_dict(state)
    model = TensorParallel(model, ctx, sequence_parallel=True).parallelize()
    cached = model.generate(input_ids=ids, max_new_tokens=5)                     # KV cache, heads sharded over the group
    assert torch.equal(cached, ref_tokens), (cached, ref_tokens)
    assert torch.equal(model.generate(input_ids=ids, max_new_tokens=5, use_cache=False), ref_tokens)
    ctx.destroy()


@pytest.mark.parametrize("tp", [2, 4])
def test_tensor_parallel_generation_with_kv_cache_matches_hf(tp):
    torch.manual_seed(3)
    model = _hf_bloom()
    ids = torch.randint(0, 96, (2, 7))
    ref = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=5, do_sample=False)
    spawn(run_tp_cached_generate, world_size=tp, tp=tp, state=copy.deepcopy(model.state_dict()), ids=ids, ref_tokens=ref)


def run_left_padded(rank, world_size, port, tp, pp, state, ids, mask, ref_loss, ref_grad):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, 1)
    model = _hf_bloom()
    model.load_state_dict(state)
    model = TensorParallel(model, ctx, sequence_parallel=True).parallelize()
    if pp > 1:
        from pipegoose_b200.nn import PipelineParallel

        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    # the reference's README loop: labels = input_ids, the tokenizer's mask passed along
    out = model(input_ids=ids, attention_mask=mask, labels=ids)
    assert torch.allclose(out.loss, ref_loss, atol=2e-5), (out.loss, ref_loss)
    out.loss.backward()
    if pp == 1:   # a tensor-parallel-replicated parameter: its partial gradients were summed over the group by the sync hook
        p = dict(model.named_parameters())["transformer.h.0.input_layernorm.weight"]
        g = p.grad if p.grad is not None else p.main_grad
        assert torch.allclose(g, ref_grad, atol=2e-5)
    ctx.destroy()


@pytest.mark.parametrize("tp,pp", [(2, 1), (2, 2)])
def test_left_padded_batches_through_the_fused_parallel_paths(tp, pp):
    torch.manual_seed(0)
    hf = _hf_bloom()
    state = copy.deepcopy(hf.state_dict())
    ids = torch.randint(1, 96, (4, 8))
    mask = torch.ones(4, 8, dtype=torch.long)
    mask[0, :3] = 0
    mask[2, :5] = 0
    ids = ids.masked_fill(mask == 0, 3)
    labels = ids.masked_fill(mask == 0, -100)
    labels[0, 3] = labels[2, 5] = -100        # see tests/test_models_bloom.py: 🤗 scores the first real token from a pad
    # (pp > 1: the pipelined loss is the token-weighted mean over micro-batches = the mean over all scored tokens)
    want = hf(input_ids=ids, attention_mask=mask, labels=labels).loss
    want.backward()
    spawn(run_left_padded, world_size=tp * pp, tp=tp, pp=pp, state=state, ids=ids, mask=mask, ref_loss=want.detach(),
          ref_grad=hf.transformer.h[0].input_layernorm.weight.grad.clone())


def test_round_trip_back_to_transformers(tmp_path):
    """fused model -> 🤗: ``to_hf()`` (same logits), ``save_hf_pretrained`` (a directory ``from_pretrained`` reads), also
    for a model that was converted in place and trained a step, and after ``deparallelize()``-style refusal of shards."""
    from transformers import BloomForCausalLM as HFBloom

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM

    torch.manual_seed(0)
    mine = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    ids = torch.randint(0, 96, (2, 7))
    hf = mine.to_hf().eval()
    assert type(hf).__module__.startswith("transformers.") and hf.lm_head.weight is hf.transformer.word_embeddings.weight
    assert torch.allclose(hf(input_ids=ids).logits, mine(ids).logits, atol=1e-5)
    mine.save_hf_pretrained(str(tmp_path / "export"))
    again = HFBloom.from_pretrained(str(tmp_path / "export")).eval()
    assert torch.allclose(again(input_ids=ids).logits, mine(ids).logits, atol=1e-5)
    # converted in place from 🤗, trained, exported: the user's own config object comes back
    start = _hf_bloom(hidden_dropout=0.0)
    cfg = start.config
    fast = convert_hf_bloom_(start)
    loss = fast(input_ids=ids, labels=ids).loss
    loss.backward()
    with torch.no_grad():
        for p in fast.parameters():
            p -= 0.1 * (p.grad if p.grad is not None else p.main_grad).to(p.dtype)
    back = fast.to_hf().eval()
    assert back.config is cfg
    assert torch.allclose(back(input_ids=ids).logits, fast.eval()(ids).logits, atol=1e-5)

    class FakeTP:
        pass

    fast.tp = FakeTP()
    with pytest.raises(ValueError, match="unsharded"):
        fast.to_hf()
