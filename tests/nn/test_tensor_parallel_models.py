"""Whole-model tensor parallelism: the 🤗 Bloom replicated-activation path (reference semantics) and
the sequence-parallel fast path of pipegoose_b200's own Bloom, both against the unsharded model."""
import copy

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import TensorParallel
from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def run_hf_bloom(rank, world_size, port, tp, state, ids, ref_logits, ref_generated):
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    model = HFBloom(HFConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    # every leaf except activations / dropout is one of the four parallel classes
    for name, mod in model.named_modules():
        if len(list(mod.children())) == 0 and len(list(mod.parameters(recurse=False))) > 0:
            assert isinstance(mod, (ColumnParallelLinear, RowParallelLinear, ParallelEmbedding, LayerNorm)), name
    assert model.lm_head.weight is model.transformer.word_embeddings.weight
    assert model.transformer.word_embeddings.weight.shape[0] == 128 // tp
    out = model(input_ids=ids, labels=ids)
    assert torch.allclose(out.logits, ref_logits, atol=1e-4)
    gen = model.generate(ids, max_new_tokens=1, do_sample=False)
    assert torch.equal(gen, ref_generated)
    out.loss.backward()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    opt.step()
    for p in model.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    ctx.destroy()


def test_hf_bloom_tensor_parallel_matches_unsharded():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    torch.manual_seed(0)
    model = HFBloom(HFConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4)).eval()
    ids = torch.randint(0, 128, (2, 8))
    with torch.no_grad():
        ref_logits = model(input_ids=ids).logits
        ref_gen = model.generate(ids, max_new_tokens=1, do_sample=False)
    spawn(run_hf_bloom, world_size=2, tp=2, state=model.state_dict(), ids=ids, ref_logits=ref_logits, ref_generated=ref_gen)


def run_fast_bloom(rank, world_size, port, tp, state, ids, ref_loss, ref_grads, ref_logits, ref_generated):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    cfg = BloomConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    r = ctx.get_local_rank(ParallelMode.TENSOR)
    blk = model.transformer.h[0]
    assert blk.self_attention.query_key_value.weight.shape == (3 * 32 // tp, 32)
    assert blk.self_attention.dense.weight.shape == (32, 32 // tp)
    assert blk.mlp.dense_h_to_4h.weight.shape == (128 // tp, 32)
    assert blk.mlp.dense_4h_to_h.weight.shape == (32, 128 // tp)
    assert model.lm_head.weight is model.transformer.word_embeddings.weight
    loss = model(ids, labels=ids).loss
    assert torch.allclose(loss, ref_loss, atol=1e-5), (loss, ref_loss)
    loss.backward()
    logits = model(ids).logits
    assert torch.allclose(logits, ref_logits, atol=1e-4)

    def shard(name, g):
        if "query_key_value" in name or "dense_h_to_4h" in name:
            return g.chunk(tp, 0)[r]
        if ("self_attention.dense.weight" in name) or ("dense_4h_to_h.weight" in name):
            return g.chunk(tp, 1)[r]
        if "word_embeddings.weight" in name:
            return g.chunk(tp, 0)[r]
        return g

    for name, p in model.named_parameters():
        # gradients of TP-replicated parameters are partial sums over the token shards; TensorParallel's
        # TensorPartialGradSync has already summed them over the tensor group when backward returned
        g = p.grad.clone()
        want = shard(name, ref_grads[name])
        assert torch.allclose(g, want, atol=2e-5), name
    # generation under token sharding: odd prompt lengths are right-padded internally (3 rows x 5, 6, 7 tokens)
    assert torch.equal(model.generate(ids[:3, :5], max_new_tokens=3), ref_generated)
    ctx.destroy()


def test_fast_bloom_sequence_parallel_matches_unsharded():
    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layernorm" in n or "ln_f" in n or "bias" in n:
                p.add_(torch.randn_like(p) * 0.1)
    ids = torch.randint(0, 128, (4, 8))
    loss = model(ids, labels=ids).loss
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    with torch.no_grad():
        logits = model(ids).logits
        generated = model.generate(ids[:3, :5], max_new_tokens=3)
    spawn(run_fast_bloom, world_size=2, tp=2, state=copy.deepcopy(model.state_dict()), ids=ids, ref_loss=loss.detach(),
          ref_grads=grads, ref_logits=logits, ref_generated=generated)


def run_tp_only_training(rank, world_size, port, state, ids, ref_losses, ref_ln_grad):
    """tp=2, dp=1: no DataParallel reducer exists, yet LayerNorm / row-parallel-bias gradients (partial sums over
    the token shards) must be summed over the TENSOR group, and the replicated parameters must stay identical."""
    import torch.distributed as dist

    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam

    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    cfg = BloomConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)

    # gradient accumulation over two micro-batches: partial sums are reduced once, after the last backward
    optim.zero_grad()
    halves = ids.chunk(2)
    with model.no_sync():
        (model(halves[0], labels=halves[0]).loss / 2).backward()
    (model(halves[1], labels=halves[1]).loss / 2).backward()
    ln = model.transformer.h[0].input_layernorm.weight
    g = ln.main_grad if getattr(ln, "main_grad", None) is not None else ln.grad
    assert torch.allclose(g.float(), ref_ln_grad, atol=2e-5), (g, ref_ln_grad)

    losses = []
    for _ in range(len(ref_losses)):
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        losses.append(loss.item())
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3, (losses, ref_losses)
    for name, p in model.named_parameters():
        if getattr(p, "tp_partial_grad", False):
            other = p.detach().clone()
            dist.broadcast(other, src=0)
            assert torch.allclose(p.detach(), other, atol=1e-7), f"{name} diverged across the tensor group"
    ctx.destroy()


def test_tensor_parallel_without_data_parallel_trains_like_a_single_process():
    from pipegoose_b200.optim import FusedAdam

    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=128, hidden_size=32, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 128, (4, 8))
    halves = ids.chunk(2)
    for h in halves:
        (model(h, labels=h).loss / 2).backward()
    ref_ln_grad = model.transformer.h[0].input_layernorm.weight.grad.clone()
    model.zero_grad()
    opt = FusedAdam(model.parameters(), lr=1e-2)
    ref_losses = []
    for _ in range(3):
        loss = model(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    spawn(run_tp_only_training, world_size=2, state=state, ids=ids, ref_losses=ref_losses, ref_ln_grad=ref_ln_grad)


def run_deparallelize(rank, world_size, port, fast, state, ids, ref_logits):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    if fast:
        model = BloomForCausalLM(BloomConfig(vocab_size=99, hidden_size=32, n_layer=2, n_head=4))
    else:
        from transformers import BloomConfig as HFConfig
        from transformers import BloomForCausalLM as HFBloom

        model = HFBloom(HFConfig(vocab_size=99, hidden_size=32, n_layer=2, n_head=4)).eval()
    model.load_state_dict(state)
    wrapper = TensorParallel(model, ctx)
    model = wrapper.parallelize()
    assert model.lm_head.weight.shape[0] < 99  # sharded
    model = wrapper.deparallelize()
    got = {k: v for k, v in model.state_dict().items()}
    for k, v in state.items():
        assert got[k].shape == v.shape and torch.equal(got[k], v), k
    assert model.lm_head.weight is model.get_input_embeddings().weight  # still tied
    with torch.no_grad():
        logits = model(ids).logits
    assert torch.allclose(logits, ref_logits, atol=1e-5)
    ctx.destroy()


@pytest.mark.parametrize("fast", [True, False])
def test_tensor_parallel_deparallelize_restores_the_model(fast):
    torch.manual_seed(0)
    if fast:
        model = BloomForCausalLM(BloomConfig(vocab_size=99, hidden_size=32, n_layer=2, n_head=4))
    else:
        from transformers import BloomConfig as HFConfig
        from transformers import BloomForCausalLM as HFBloom

        model = HFBloom(HFConfig(vocab_size=99, hidden_size=32, n_layer=2, n_head=4)).eval()
    ids = torch.randint(0, 99, (2, 8))
    with torch.no_grad():
        ref = model(ids).logits
    spawn(run_deparallelize, world_size=2, fast=fast, state=copy.deepcopy(model.state_dict()), ids=ids, ref_logits=ref)


# ---------------------------------------------------------------- vocabulary sizes that do not split evenly
def run_odd_vocab_fast(rank, world_size, port, state, ids, ref_loss, ref_logits, ref_losses):
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam

    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=101, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    assert model.lm_head.weight.shape[0] * 2 > 101                      # zero-padded table
    assert torch.allclose(model(ids, labels=ids).loss, ref_loss, atol=1e-5)   # phantom classes are not in the softmax
    logits = model(ids).logits
    assert logits.shape == ref_logits.shape and torch.allclose(logits, ref_logits, atol=1e-4)
    opt = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    for want in ref_losses:
        loss = model(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(loss.item() - want) < 1e-4, (loss.item(), ref_losses)
    ctx.destroy()


def test_fast_bloom_with_a_vocabulary_that_needs_padding():
    from pipegoose_b200.optim import FusedAdam

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=101, hidden_size=32, n_layer=2, n_head=4))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 101, (2, 8))
    ref_loss, ref_logits = model(ids, labels=ids).loss.detach(), model(ids).logits.detach()
    opt, ref_losses = FusedAdam(model.parameters(), lr=1e-2), []
    for _ in range(3):
        loss = model(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    spawn(run_odd_vocab_fast, world_size=2, state=state, ids=ids, ref_loss=ref_loss, ref_logits=ref_logits,
          ref_losses=ref_losses)


def run_odd_vocab_hf(rank, world_size, port, state, ids, ref_loss, ref_logits):
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    model = HFBloom(HFConfig(vocab_size=101, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    wrapper = TensorParallel(model, ctx)
    model = wrapper.parallelize()
    out = model(input_ids=ids, labels=ids)
    assert out.logits.shape == ref_logits.shape and torch.allclose(out.logits, ref_logits, atol=1e-4)
    assert torch.allclose(out.loss, ref_loss, atol=1e-5)
    out.loss.backward()
    model = wrapper.deparallelize()
    assert model.lm_head.weight.shape == (101, 32) and torch.allclose(model(input_ids=ids).logits, ref_logits, atol=1e-4)
    ctx.destroy()


def test_hf_bloom_with_a_vocabulary_that_needs_padding():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    torch.manual_seed(0)
    model = HFBloom(HFConfig(vocab_size=101, hidden_size=32, n_layer=2, n_head=4)).eval()
    ids = torch.randint(0, 101, (2, 8))
    out = model(input_ids=ids, labels=ids)
    spawn(run_odd_vocab_hf, world_size=2, state=copy.deepcopy(model.state_dict()), ids=ids, ref_loss=out.loss.detach(),
          ref_logits=out.logits.detach())
