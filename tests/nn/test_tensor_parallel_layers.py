"""Reference-compatible tensor-parallel layers vs their unsharded counterparts (outputs AND grads),
following the reference's strategy: reference values are computed in the parent and compared
inside spawned gloo ranks (tests/nn/tensor_parallel/test_linear.py:131-173 in the reference)."""
import pytest
import torch
from torch import nn

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear
from pipegoose_b200.nn.tensor_parallel.loss import VocabParallelCrossEntropy
from pipegoose_b200.testing.utils import get_partition, init_parallel_context, spawn


def run_layers(rank, world_size, port, tp, x, w_col, b_col, w_row, b_row, ref):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    # ---- column parallel (gather_output=True)
    col = ColumnParallelLinear(x.shape[-1], w_col.shape[0], bias=True, gather_output=True, parallel_context=ctx)
    col.weight.data = get_partition(w_col, 0, ctx).clone()
    col.bias.data = get_partition(b_col, 0, ctx).clone()
    xin = x.clone().requires_grad_(True)
    out = col(xin)
    assert torch.allclose(out, ref["col_out"], atol=1e-5)
    out.sum().backward()
    assert torch.allclose(xin.grad, ref["col_dx"], atol=1e-5)
    assert torch.allclose(col.weight.grad, get_partition(ref["col_dw"], 0, ctx), atol=1e-5)
    assert torch.allclose(col.bias.grad, get_partition(ref["col_db"], 0, ctx), atol=1e-5)
    # ---- row parallel
    row = RowParallelLinear(w_row.shape[1], w_row.shape[0], bias=True, parallel_context=ctx)
    row.weight.data = get_partition(w_row, 1, ctx).clone()
    row.bias.data = b_row.clone()
    xin = x.clone().requires_grad_(True)
    out = row(xin)
    assert torch.allclose(out, ref["row_out"], atol=1e-5)
    out.sum().backward()
    assert torch.allclose(xin.grad, ref["row_dx"], atol=1e-5)
    assert torch.allclose(row.weight.grad, get_partition(ref["row_dw"], 1, ctx), atol=1e-5)
    assert torch.allclose(row.bias.grad, ref["row_db"], atol=1e-5)
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_column_and_row_parallel_linear(tp):
    torch.manual_seed(0)
    x = torch.randn(5, 16)
    col, row = nn.Linear(16, 8), nn.Linear(16, 8)
    ref = {}
    for name, lin in (("col", col), ("row", row)):
        xin = x.clone().requires_grad_(True)
        out = lin(xin)
        out.sum().backward()
        ref.update({f"{name}_out": out.detach(), f"{name}_dx": xin.grad, f"{name}_dw": lin.weight.grad, f"{name}_db": lin.bias.grad})
    spawn(run_layers, world_size=tp, tp=tp, x=x, w_col=col.weight.data, b_col=col.bias.data,
          w_row=row.weight.data, b_row=row.bias.data, ref=ref)


def run_embedding_ln_loss(rank, world_size, port, tp, ids, table, ref_emb, ref_demb, logits, targets, ref_loss, ref_dlogits):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    emb = ParallelEmbedding(table.shape[0], table.shape[1], ctx)
    emb.weight.data = get_partition(table, 0, ctx).clone()
    out = emb(ids)
    assert torch.allclose(out, ref_emb, atol=1e-6)
    out.sum().backward()
    assert torch.allclose(emb.weight.grad, get_partition(ref_demb, 0, ctx), atol=1e-6)
    assert emb.vocab_end_idx - emb.vocab_start_idx == table.shape[0] // tp
    # layer norm is replicated
    ln = LayerNorm(table.shape[1], parallel_context=ctx)
    assert torch.allclose(ln(out.detach()), torch.nn.functional.layer_norm(out.detach(), (table.shape[1],)), atol=1e-6)
    # vocab-parallel cross entropy: forward and the (fixed) backward
    local = get_partition(logits, -1, ctx).clone().requires_grad_(True)
    loss = VocabParallelCrossEntropy(ctx)(local, targets)
    assert torch.allclose(loss, ref_loss, atol=1e-5)
    loss.backward()
    assert torch.allclose(local.grad, get_partition(ref_dlogits, -1, ctx), atol=1e-6)
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_embedding_layernorm_and_vocab_parallel_loss(tp):
    torch.manual_seed(1)
    table = torch.randn(12, 6)
    ids = torch.randint(0, 12, (2, 5))
    t = table.clone().requires_grad_(True)
    ref_emb = torch.nn.functional.embedding(ids, t)
    ref_emb.sum().backward()
    logits = torch.randn(2, 5, 12)
    targets = torch.randint(0, 12, (2, 5))
    lg = logits.clone().requires_grad_(True)
    ref_loss = torch.nn.functional.cross_entropy(lg.view(-1, 12), targets.view(-1))
    ref_loss.backward()
    spawn(run_embedding_ln_loss, world_size=tp, tp=tp, ids=ids, table=table, ref_emb=ref_emb.detach(), ref_demb=t.grad,
          logits=logits, targets=targets, ref_loss=ref_loss.detach(), ref_dlogits=lg.grad)


def test_parallel_mapping():
    from pipegoose_b200.nn.tensor_parallel.parallel_mapping import TensorParallelMapping as M

    assert M.is_column_parallel("transformer.h.0.mlp.dense_h_to_4h")
    assert M.is_column_parallel("transformer.h.3.self_attention.query_key_value")
    assert M.is_row_parallel("transformer.h.0.mlp.dense_4h_to_h")
    assert M.is_row_parallel("transformer.h.0.self_attention.dense")
    assert M.is_lm_head("lm_head")
    assert not M.is_column_parallel("transformer.h.0.self_attention.dense")
    assert not M.is_row_parallel("transformer.word_embeddings")


def test_vocab_utility():
    from pipegoose_b200.nn.tensor_parallel._utils import VocabUtility

    assert VocabUtility.get_vocab_range_idx_from_partition_size(10, 2) == (20, 30)          # (partition_size, rank)
    assert VocabUtility.get_vocab_range_from_global_vocab_size(4, 3, 40) == (30, 40)          # (world_size, rank, vocab_size)
    assert VocabUtility.get_vocab_range_from_per_partition_vocab_size(10, 2) == (20, 30)
