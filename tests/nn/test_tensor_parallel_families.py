"""Class-swap tensor parallelism on more 🤗 families than the reference's mapping lists (OPT, LLaMA, GPT-NeoX, BERT,
GPT-2) and on a user-registered mapping: logits, loss and the gradients of the sharded layers match the unsharded model."""
import copy

import pytest
import torch
from torch import nn

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn import TensorParallel
from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear
from pipegoose_b200.nn.tensor_parallel.parallel_mapping import Column, LMHead, Row, TensorParallelMapping
from pipegoose_b200.testing.utils import init_parallel_context, spawn

V, H = 96, 32


def build(family):
    import transformers as T

    if family == "opt":
        cfg = T.OPTConfig(vocab_size=V, hidden_size=H, num_hidden_layers=2, ffn_dim=64, num_attention_heads=4,
                          max_position_embeddings=32, word_embed_proj_dim=H, dropout=0.0, attention_dropout=0.0)
        return T.OPTForCausalLM(cfg)
    if family == "llama":
        cfg = T.LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                            num_key_value_heads=2, max_position_embeddings=32, tie_word_embeddings=False)
        return T.LlamaForCausalLM(cfg)
    if family == "gpt_neox":
        cfg = T.GPTNeoXConfig(vocab_size=V, hidden_size=H, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                              max_position_embeddings=32, hidden_dropout=0.0, attention_dropout=0.0)
        return T.GPTNeoXForCausalLM(cfg)
    if family == "gpt2":
        cfg = T.GPT2Config(vocab_size=V, n_embd=H, n_layer=2, n_head=4, n_positions=32, resid_pdrop=0.0, embd_pdrop=0.0,
                           attn_pdrop=0.0)
        return T.GPT2LMHeadModel(cfg)
    if family == "bert":
        cfg = T.BertConfig(vocab_size=V, hidden_size=H, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                           max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        return T.BertForMaskedLM(cfg)
    raise ValueError(family)


EXPECT = {  # family -> (column-parallel leaves, row-parallel leaves, vocab-parallel embeddings) per model
    "opt": (2 * 4 + 1, 2 * 2, 1), "llama": (2 * 5 + 1, 2 * 2, 1), "gpt_neox": (2 * 2 + 1, 2 * 2, 1),
    "gpt2": (1, 0, 1), "bert": (2 * 4 + 1, 2 * 2, 1),
}


def run_family(rank, world_size, port, family, state, ids, ref_logits, ref_loss, ref_grads):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    model = build(family)
    model.load_state_dict(state)
    model.train()
    names = {id(m): n for n, m in model.named_modules()}
    model = TensorParallel(model, ctx).parallelize()
    kinds = [type(m) for m in model.modules()]
    counts = tuple(sum(k is c for k in kinds) for c in (ColumnParallelLinear, RowParallelLinear, ParallelEmbedding))
    assert counts == EXPECT[family], (family, counts)
    emb, head = model.get_input_embeddings(), model.get_output_embeddings()
    if head is not None and getattr(model.config, "tie_word_embeddings", False):
        assert head.weight is emb.weight, "the tie must survive the sharding"
    out = model(input_ids=ids, labels=ids)
    assert torch.allclose(out.logits, ref_logits, atol=2e-4), family
    assert torch.allclose(out.loss, ref_loss, atol=1e-5)
    out.loss.backward()
    r = ctx.get_local_rank(ParallelMode.TENSOR)
    checked = 0
    for m in model.modules():
        if isinstance(m, (ColumnParallelLinear, RowParallelLinear)) and not getattr(m, "_pg_tied_to_embedding", False):
            full = ref_grads.get(names[id(m)] + ".weight")
            if full is None:
                continue
            dim = 0 if isinstance(m, ColumnParallelLinear) else 1
            if full.shape[dim] % 2:
                continue
            want = full.chunk(2, dim=dim)[r]
            assert torch.allclose(m.weight.grad, want, atol=2e-5), names[id(m)]
            checked += 1
    assert checked > 0 or family == "gpt2"
    ctx.destroy()


@pytest.mark.parametrize("family", ["opt", "llama", "gpt_neox", "gpt2", "bert"])
def test_hf_families_tensor_parallel(family):
    torch.manual_seed(0)
    model = build(family)
    model.train()
    ids = torch.randint(3, V, (2, 8))
    out = model(input_ids=ids, labels=ids)
    out.loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    spawn(run_family, world_size=2, family=family, state=copy.deepcopy(model.state_dict()), ids=ids,
          ref_logits=out.logits.detach(), ref_loss=out.loss.detach(), ref_grads=grads)


class TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.tok = nn.Embedding(V, H)
        self.blocks = nn.ModuleList([nn.ModuleDict({"attn": nn.ModuleDict({"wqkv": nn.Linear(H, 3 * H), "wo": nn.Linear(H, H)})})
                                     for _ in range(2)])
        self.head = nn.Linear(H, V, bias=False)

    def forward(self, ids):
        x = self.tok(ids)
        for b in self.blocks:
            x = x + b["attn"]["wo"](torch.tanh(b["attn"]["wqkv"](x))[..., :H])
        return self.head(x)


def run_registered(rank, world_size, port, state, ids, ref):
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 1)
    TensorParallelMapping.register("tiny-net", [Column(("attn.wqkv",)), Row(("attn.wo",)), LMHead(("head",))])
    try:
        model = TinyNet()
        model.load_state_dict(state)
        model = TensorParallel(model, ctx).parallelize()
        assert isinstance(model.blocks[0]["attn"]["wqkv"], ColumnParallelLinear)
        assert isinstance(model.blocks[1]["attn"]["wo"], RowParallelLinear)
        assert model.head.weight.shape[0] == V // 2
        assert torch.allclose(model(ids), ref, atol=1e-5)
    finally:
        TensorParallelMapping.unregister("tiny-net")
    assert not TensorParallelMapping.is_column_parallel("blocks.0.attn.wqkv")
    ctx.destroy()


def test_user_registered_mapping():
    torch.manual_seed(0)
    model = TinyNet()
    ids = torch.randint(0, V, (2, 8))
    spawn(run_registered, world_size=2, state=copy.deepcopy(model.state_dict()), ids=ids, ref=model(ids).detach())
