"""pipegoose_b200.models.gpt2: the GPT-2 family on the Bloom blocks (zero ALiBi slopes, learned positions, no embedding
LayerNorm) — against 🤗 GPT-2, and through the tensor / pipeline / data parallel wrappers against the unsharded model."""
import copy

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel
from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=4, n_head=4, n_positions=16)


def _hf():
    from transformers import GPT2Config as HFConfig
    from transformers import GPT2LMHeadModel as HFModel

    return HFModel(HFConfig(vocab_size=96, n_embd=32, n_layer=2, n_head=4, n_positions=16, resid_pdrop=0.0, embd_pdrop=0.0,
                            attn_pdrop=0.0))


def test_matches_hf_gpt2_and_round_trips_the_state_dict():
    torch.manual_seed(0)
    hf = _hf().eval()
    for p in hf.parameters():  # 🤗 initialises biases to zero: make them count
        if p.dim() == 1:
            p.data.normal_(std=0.05)
    mine = GPT2LMHeadModel.from_hf(hf)
    ids = torch.randint(0, 96, (2, 12))
    want = hf(input_ids=ids, labels=ids)
    got = mine(ids, labels=ids)
    assert torch.allclose(got.loss, want.loss, atol=1e-5)
    assert torch.allclose(mine(ids).logits, want.logits, atol=1e-4)
    got.loss.backward()
    want.loss.backward()
    hf_grads = GPT2LMHeadModel.convert_hf_state_dict({n: p.grad for n, p in hf.named_parameters()}, 4)
    for n, p in mine.named_parameters():
        assert torch.allclose(p.grad, hf_grads[n], atol=2e-5), n
    back = mine.to_hf_state_dict()
    for n, t in hf.state_dict().items():
        if n.endswith(".attn.bias") or n.endswith(".attn.masked_bias"):
            continue
        assert torch.equal(back[n], t), n
    assert torch.equal(mine.generate(ids, max_new_tokens=2), hf.generate(ids, max_new_tokens=2, do_sample=False))


def run_layout(rank, world_size, port, tp, pp, dp, state, ids, ref_losses):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model = GPT2LMHeadModel(GPT2Config(**CFG))
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    losses = []
    for _ in range(len(ref_losses)):
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        losses.append(loss.item())
    t = torch.tensor(losses)
    dist.all_reduce(t)
    mean = (t / world_size).tolist()
    for a, b in zip(mean, ref_losses):
        assert abs(a - b) < 2e-3, (mean, ref_losses)
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,dp", [(2, 1, 1), (2, 1, 2), (2, 2, 1)])
def test_parallel_layouts_follow_single_process_training(tp, pp, dp):
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(**CFG))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    opt = FusedAdam(model.parameters(), lr=1e-2)
    chunks = [mb for rep in ids.chunk(dp) for mb in (rep.chunk(2) if pp > 1 else [rep])]
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        total = 0.0
        for mb in chunks:
            loss = model(mb, labels=mb).loss / len(chunks)
            loss.backward()
            total += loss.item()
        opt.step()
        ref_losses.append(total)
    assert ref_losses[-1] < ref_losses[0]
    spawn(run_layout, world_size=tp * pp * dp, tp=tp, pp=pp, dp=dp, state=state, ids=ids, ref_losses=ref_losses)


def test_gpt2_generate_left_padded_prompts_match_transformers():
    """Learned absolute positions count from each row's first real token (KV-cache path: ``_learned_positions``; training
    forward: rows rotated to right padding)."""
    from transformers import GPT2Config as HFConfig
    from transformers import GPT2LMHeadModel as HFGPT2

    from pipegoose_b200.models.gpt2 import GPT2LMHeadModel

    torch.manual_seed(0)
    hf = HFGPT2(HFConfig(vocab_size=96, n_embd=32, n_layer=2, n_head=4, n_positions=64, resid_pdrop=0.0, embd_pdrop=0.0,
                         attn_pdrop=0.0)).eval()
    mine = GPT2LMHeadModel.from_hf(hf).eval()
    ids = torch.randint(1, 96, (3, 7), generator=torch.Generator().manual_seed(1))
    mask = torch.ones_like(ids)
    mask[0, :3] = 0
    mask[2, :5] = 0
    ids = ids * mask
    want = hf.generate(input_ids=ids, attention_mask=mask, max_new_tokens=5, do_sample=False, pad_token_id=0)
    for use_cache in (True, False):
        assert torch.equal(mine.generate(ids, attention_mask=mask, max_new_tokens=5, use_cache=use_cache), want)


def test_gpt2_round_trip_back_to_transformers(tmp_path):
    from transformers import GPT2LMHeadModel as HFGPT2

    from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(0)
    mine = GPT2LMHeadModel(GPT2Config.gpt2_tiny()).eval()
    ids = torch.randint(0, mine.config.vocab_size, (2, 9))
    hf = mine.to_hf().eval()
    assert torch.allclose(hf(input_ids=ids).logits, mine(ids).logits, atol=1e-5)
    mine.save_hf_pretrained(str(tmp_path / "export"))
    again = HFGPT2.from_pretrained(str(tmp_path / "export")).eval()
    assert torch.allclose(again(input_ids=ids).logits, mine(ids).logits, atol=1e-5)
    back = GPT2LMHeadModel.from_hf(again).eval()
    assert torch.allclose(back(ids).logits, mine(ids).logits, atol=1e-5)
