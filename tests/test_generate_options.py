"""``generate()`` beyond greedy decoding of equal-length prompts: left-padded prompts of unequal length (what 🤗's Bloom
tokenizer produces for a batch; the mask used to be ignored — pads were read as tokens), end-of-sequence handling, and
sampling (temperature / top-k / top-p), identical on every tensor-parallel rank."""
import copy

import pytest
import torch

from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def _pair():
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    torch.manual_seed(0)
    hf = HFBloom(HFConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)).eval()
    return hf, BloomForCausalLM.from_hf(hf).eval()


def test_left_padded_prompts_of_unequal_length_match_transformers():
    hf, mine = _pair()
    ids = torch.randint(1, 96, (3, 7), generator=torch.Generator().manual_seed(1))
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    mask[2, :5] = 0
    ids = ids * mask                                           # pad id 0 in front
    want = hf.generate(input_ids=ids, attention_mask=mask, max_new_tokens=5, do_sample=False, pad_token_id=0)
    calls = []
    inner = mine._incremental_logits
    mine._incremental_logits = lambda *a, **k: (calls.append(a[0].shape[1]), inner(*a, **k))[1]
    got = mine.generate(ids, attention_mask=mask, max_new_tokens=5)
    del mine._incremental_logits
    assert torch.equal(got, want)
    assert calls == [7, 1, 1, 1, 1]                            # KV cache: prompt once, then one position per token
    assert torch.equal(mine.generate(ids, attention_mask=mask, max_new_tokens=5, use_cache=False), want)
    # a mask that is not left-padded (pads behind the prompt) is refused: the new tokens would follow pads
    right = torch.ones_like(ids)
    right[1, 5:] = 0
    with pytest.raises(ValueError, match="LEFT-padded"):
        mine.generate(ids * right, attention_mask=right, max_new_tokens=1)
    # every row equals what it generates alone, without padding
    for r, skip in ((0, 3), (2, 5)):
        alone = mine.generate(ids[r:r + 1, skip:], max_new_tokens=5)
        assert torch.equal(alone[0, -5:], got[r, -5:])
    # a mask without pads takes the cached path
    full = torch.ones_like(ids)
    assert torch.equal(mine.generate(ids, attention_mask=full, max_new_tokens=4), mine.generate(ids, max_new_tokens=4))


def test_end_of_sequence_handling_matches_transformers():
    hf, mine = _pair()
    ids = torch.randint(1, 96, (4, 5), generator=torch.Generator().manual_seed(2))
    greedy = mine.generate(ids, max_new_tokens=8)
    eos = int(greedy[0, 6])                                    # a token row 0 emits as its second new token
    want = hf.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=8, do_sample=False,
                       eos_token_id=eos, pad_token_id=0)
    for use_cache in (True, False):
        got = mine.generate(ids, max_new_tokens=8, eos_token_id=eos, pad_token_id=0, use_cache=use_cache)
        assert torch.equal(got, want)
    first = int((got[0, 5:] == eos).float().argmax()) + 5      # row 0's first eos among the new tokens
    assert first <= 6 and (got[0, first + 1:] == 0).all()
    # all rows finished -> the loop stops early
    one = mine.generate(ids[:1], max_new_tokens=8, eos_token_id=eos)
    assert one.shape[1] == first + 1 and one[0, -1] == eos


def test_sampling_filters_and_reproducibility():
    _, mine = _pair()
    ids = torch.randint(1, 96, (2, 6), generator=torch.Generator().manual_seed(3))
    greedy = mine.generate(ids, max_new_tokens=6)
    assert torch.equal(mine.generate(ids, max_new_tokens=6, do_sample=True, top_k=1), greedy)
    assert torch.equal(mine.generate(ids, max_new_tokens=6, do_sample=True, top_p=1e-6), greedy)
    assert torch.equal(mine.generate(ids, max_new_tokens=6, do_sample=True, temperature=1e-4), greedy)
    a = mine.generate(ids, max_new_tokens=6, do_sample=True, temperature=1.5, generator=torch.Generator().manual_seed(7))
    b = mine.generate(ids, max_new_tokens=6, do_sample=True, temperature=1.5, generator=torch.Generator().manual_seed(7))
    assert torch.equal(a, b) and not torch.equal(a, greedy)
    # the filters: a sampled token is always among the k most likely / inside the nucleus
    logits = torch.randn(64, 96, generator=torch.Generator().manual_seed(4)) * 3
    g = torch.Generator().manual_seed(5)
    for _ in range(5):
        tok = mine._select_token(logits, True, 1.0, 4, 1.0, g)
        assert (logits.topk(4, -1).indices == tok[:, None]).any(-1).all()
        tok = mine._select_token(logits, True, 1.0, 0, 0.5, g)
        probs = logits.softmax(-1)
        sorted_p, order = probs.sort(-1, descending=True)
        before = sorted_p.cumsum(-1) - sorted_p
        rank_of = (order == tok[:, None]).float().argmax(-1)
        assert (before.gather(1, rank_of[:, None]) < 0.5).all()
    # frequencies follow the distribution
    row = torch.tensor([[2.0, 1.0, 0.0, -1.0] + [-30.0] * 92])
    draws = torch.stack([mine._select_token(row, True, 1.0, 0, 1.0, g) for _ in range(2000)]).squeeze(1)
    freq = torch.bincount(draws, minlength=96)[:4].float() / 2000
    assert torch.allclose(freq, row.softmax(-1)[0, :4], atol=0.04)


def run_tp_sampling(rank, world_size, port, state, ids, mask, want):
    import torch.distributed as dist

    from pipegoose_b200.nn import TensorParallel

    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize().eval()
    torch.manual_seed(100 + rank)                              # different RNG streams: the group must still agree
    sampled = model.generate(ids, max_new_tokens=6, do_sample=True, temperature=1.3, top_k=20)
    ragged = model.generate(ids, attention_mask=mask, max_new_tokens=4)          # KV cache, heads sharded, pad keys masked
    assert torch.equal(ragged, want), (ragged, want)
    assert torch.equal(model.generate(ids, attention_mask=mask, max_new_tokens=4, use_cache=False), want)
    both = [None, None]
    dist.all_gather_object(both, (sampled.tolist(), ragged.tolist()))
    assert both[0] == both[1]
    ctx.destroy()
    return None


def test_tensor_parallel_ranks_agree_on_sampled_tokens_and_ragged_prompts():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)).eval()
    ids = torch.randint(1, 96, (2, 5), generator=torch.Generator().manual_seed(6))
    mask = torch.ones_like(ids)
    mask[1, :2] = 0
    want = model.generate(ids * mask, attention_mask=mask, max_new_tokens=4)
    spawn(run_tp_sampling, world_size=2, state=copy.deepcopy(model.state_dict()), ids=ids * mask, mask=mask, want=want)
    assert want.shape == (2, 9)


def run_pipelined_generate(rank, world_size, port, tp, state, ids, mask, want, want_ragged, eos):
    import torch.distributed as dist

    from pipegoose_b200.nn import PipelineParallel, TensorParallel

    ctx = init_parallel_context(rank, world_size, port, tp, 2, 1)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize().eval()
    assert torch.equal(model.generate(ids, max_new_tokens=5), want)                    # one forward-only schedule per token
    got = model.generate(ids * mask, attention_mask=mask, max_new_tokens=5, eos_token_id=eos, pad_token_id=0)
    assert torch.equal(got, want_ragged), (got, want_ragged)
    torch.manual_seed(100 + rank)                                                      # ranks draw differently, must agree
    sampled = model.generate(ids, max_new_tokens=5, do_sample=True, temperature=1.3, top_k=20)
    everyone = [None] * world_size
    dist.all_gather_object(everyone, sampled.tolist())
    assert all(e == everyone[0] for e in everyone)
    # training still works afterwards (the engine's forward-only runs left no state behind)
    model.train()
    loss = model(ids, labels=ids).loss
    loss.backward()
    assert torch.isfinite(loss)
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_generate_on_a_pipelined_model(tp):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=4, n_head=4)).eval()
    ids = torch.randint(1, 96, (4, 6), generator=torch.Generator().manual_seed(8))
    mask = torch.ones_like(ids)
    mask[1, :2] = 0
    mask[3, :4] = 0
    want = model.generate(ids, max_new_tokens=5)
    eos = int(model.generate(ids * mask, attention_mask=mask, max_new_tokens=5)[0, 8])
    want_ragged = model.generate(ids * mask, attention_mask=mask, max_new_tokens=5, eos_token_id=eos, pad_token_id=0)
    spawn(run_pipelined_generate, world_size=2 * tp, tp=tp, state=copy.deepcopy(model.state_dict()), ids=ids, mask=mask,
          want=want, want_ragged=want_ragged, eos=eos)
