"""``pipegoose_b200.models.bloom`` against 🤗 transformers' Bloom (random init, no downloads): same module tree and
parameter names, same logits / loss / gradients / greedy generation."""
import pytest
import torch

from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM


def _hf():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    torch.manual_seed(0)
    return HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)).eval()


def test_from_hf_matches_transformers_bloom():
    hf = _hf()
    model = BloomForCausalLM.from_hf(hf).eval()
    # identical parameter names / shapes: state dicts load in both directions
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in hf.state_dict().items() if k in model.state_dict()}
    hf.load_state_dict(model.state_dict(), strict=False)
    ids = torch.randint(0, 96, (2, 8))
    a = hf(input_ids=ids, labels=ids)
    b = model(ids, labels=ids)
    assert torch.allclose(a.loss, b.loss, atol=1e-5)
    a.loss.backward()
    b.loss.backward()
    hf_grads = dict(hf.named_parameters())
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, hf_grads[n].grad, atol=2e-5), n
    with torch.no_grad():
        assert torch.allclose(hf(input_ids=ids).logits, model(ids).logits, atol=1e-5)
        gen_hf = hf.generate(ids, max_new_tokens=3, do_sample=False)
        gen = model.generate(ids, max_new_tokens=3)
    assert torch.equal(gen, gen_hf)


def test_config_presets_and_flops():
    assert BloomConfig.bloom_560m().hidden_size == 1024 and BloomConfig.bloom_560m().n_layer == 24
    assert BloomConfig.bloom_3b().hidden_size // BloomConfig.bloom_3b().n_head == 80
    assert BloomConfig.bloom_7b1().n_head == 32
    m = BloomForCausalLM(BloomConfig(vocab_size=64, hidden_size=16, n_layer=1, n_head=2))
    assert m.lm_head.weight is m.transformer.word_embeddings.weight  # tied
    assert m.num_parameters() == sum(p.numel() for p in set(m.parameters()))
    assert m.flops_per_token(128) > 6 * (12 * 16 * 16 + 64 * 16)


def test_block_recompute_gives_identical_loss_and_gradients():
    """config.recompute="block": same numbers, fewer saved activations."""
    import copy

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM

    torch.manual_seed(0)
    base = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=3, n_head=4))
    ckpt = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=3, n_head=4, recompute="block"))
    ckpt.load_state_dict(copy.deepcopy(base.state_dict()))
    ids = torch.randint(0, 96, (2, 8))
    saved = {}
    for name, model in (("base", base), ("ckpt", ckpt)):
        count = [0]

        def pack(t, count=count):
            count[0] += t.numel()
            return t

        with torch.autograd.graph.saved_tensors_hooks(pack, lambda t: t):
            loss = model(ids, labels=ids).loss
        saved[name] = count[0]
        loss.backward()
    assert torch.equal(base(ids, labels=ids).loss, ckpt(ids, labels=ids).loss)
    for (n, a), (_, b) in zip(base.named_parameters(), ckpt.named_parameters()):
        assert torch.allclose(a.grad, b.grad, atol=1e-6), n
    assert saved["ckpt"] < 0.6 * saved["base"], saved


@pytest.mark.parametrize("family", ["bloom", "gpt2"])
def test_cached_generation_equals_full_recompute(family):
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(0)
    if family == "bloom":
        model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=3, n_head=4))
    else:
        model = GPT2LMHeadModel(GPT2Config(vocab_size=96, hidden_size=32, n_layer=3, n_head=4, n_positions=32))
    for p in model.parameters():
        if p.dim() == 1:
            p.data.normal_(std=0.3)   # biases / LayerNorm parameters that matter
    ids = torch.randint(0, 96, (3, 7))
    fast = model.generate(ids, max_new_tokens=6)
    slow = model.generate(ids, max_new_tokens=6, use_cache=False)
    assert fast.shape == (3, 13) and torch.equal(fast, slow)
    assert torch.equal(model.generate(ids, max_new_tokens=1), slow[:, :8])


def test_right_padding_is_harmless_and_masks_with_holes_are_refused():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    from pipegoose_b200.models.bloom import BloomForCausalLM

    torch.manual_seed(0)
    hf = HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)).eval()
    mine = BloomForCausalLM.from_hf(hf)
    ids = torch.randint(0, 96, (2, 8))
    mask = torch.ones(2, 8, dtype=torch.long)
    mask[1, 5:] = 0                                   # right padding on the second sequence
    labels = ids.masked_fill(mask == 0, -100)
    want = hf(input_ids=ids, attention_mask=mask, labels=labels).loss
    got = mine(ids, attention_mask=mask, labels=labels).loss
    assert torch.allclose(got, want, atol=1e-5)
    holes = mask.clone()
    holes[0, 3] = 0                                   # a pad between real tokens: neither left nor right padding
    with pytest.raises((RuntimeError, AssertionError)):
        mine(ids, attention_mask=holes, labels=labels)


def test_left_padded_batches_match_transformers():
    """What 🤗's Bloom tokenizer produces (padding_side = "left") and the reference's README loop feeds to the model."""
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    from pipegoose_b200.models.bloom import BloomForCausalLM, left_align

    torch.manual_seed(0)
    hf = HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)).eval()
    mine = BloomForCausalLM.from_hf(hf)
    ids = torch.randint(1, 96, (3, 9))
    mask = torch.ones(3, 9, dtype=torch.long)
    mask[0, :4] = 0                                   # left padding, different amounts per row; row 2 has none
    mask[1, :1] = 0
    ids = ids.masked_fill(mask == 0, 3)               # the pad token
    idx, keep, inv = left_align(mask)
    assert idx[0].tolist() == [4, 5, 6, 7, 8, 0, 1, 2, 3] and keep[0].tolist() == [True] * 5 + [False] * 4
    assert torch.equal(ids.gather(1, idx).gather(1, inv), ids)
    labels = ids.masked_fill(mask == 0, -100)
    # 🤗 also scores the first real token of a left-padded row as a prediction made FROM the last pad position; here a
    # row starts with its first real token, which has nothing to be predicted from — take that target out on the 🤗 side
    labels_hf = labels.clone()
    labels_hf[0, 4] = labels_hf[1, 1] = -100
    want = hf(input_ids=ids, attention_mask=mask, labels=labels_hf)
    got = mine(ids, attention_mask=mask, labels=labels).loss
    assert torch.allclose(got, want.loss, atol=1e-5), (got, want.loss)
    # pads are never scored, also when the caller leaves them in the labels (labels = input_ids, as in the README loop)
    assert torch.allclose(mine(ids, attention_mask=mask, labels=ids).loss, want.loss, atol=1e-5)
    labels = labels_hf
    logits = mine(ids, attention_mask=mask).logits    # in the caller's layout
    real = mask.bool()
    assert torch.allclose(logits[real], want.logits[real], atol=1e-4)
    # gradients flow to the same places
    hf.zero_grad(); mine.zero_grad()
    hf(input_ids=ids, attention_mask=mask, labels=labels).loss.backward()
    mine(ids, attention_mask=mask, labels=labels).loss.backward()
    g_hf = hf.transformer.h[0].self_attention.query_key_value.weight.grad
    g_me = mine.transformer.h[0].self_attention.query_key_value.weight.grad
    assert torch.allclose(g_me, g_hf, atol=1e-5)
