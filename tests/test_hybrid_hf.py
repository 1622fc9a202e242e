"""The reference's end-to-end acceptance test (tests/test_hybrid.py:19-78) on a random-init 🤗 Bloom: TensorParallel
(class-swap path) x DataParallel + DistributedOptimizer(torch.optim.Adam), one optimizer step, every parameter
must equal the matching partition of the single-process model after the same step."""
import copy

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn import DataParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer
from pipegoose_b200.testing.utils import init_parallel_context, spawn


def _hf_bloom():
    from transformers import BloomConfig, BloomForCausalLM

    return BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))


def _partition_of(name, full, tp, r):
    if "query_key_value" in name or "dense_h_to_4h" in name:
        return full.chunk(tp, 0)[r]
    if name.endswith("self_attention.dense.weight") or name.endswith("dense_4h_to_h.weight"):
        return full.chunk(tp, 1)[r]
    if "word_embeddings.weight" in name or name == "lm_head.weight":
        return full.chunk(tp, 0)[r]
    return full


def run_hybrid(rank, world_size, port, tp, dp, state, ids, ref_loss, ref_params):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    model = _hf_bloom()
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    optim = DistributedOptimizer(torch.optim.Adam(model.parameters(), lr=1e-3), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    out = model(input_ids=local, attention_mask=torch.ones_like(local), labels=local)
    optim.zero_grad()
    out.loss.backward()
    optim.step()
    # data-parallel mean of the replica losses == full-batch loss (equal shard sizes)
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


@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2), (4, 1)])
def test_hf_bloom_tensor_x_data_parallel_one_adam_step(tp, dp):
    torch.manual_seed(0)
    model = _hf_bloom()
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, 96, (4, 8))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss = model(input_ids=ids, attention_mask=torch.ones_like(ids), labels=ids).loss
    opt.zero_grad()
    loss.backward()
    opt.step()
    ref_params = {n: p.detach().clone() for n, p in model.named_parameters()}
    ref_params["lm_head.weight"] = ref_params["transformer.word_embeddings.weight"]
    spawn(run_hybrid, world_size=tp * dp, tp=tp, dp=dp, state=state, ids=ids, ref_loss=loss.detach(), ref_params=ref_params)
