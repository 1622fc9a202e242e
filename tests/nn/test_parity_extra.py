"""Unit-level parity with the reference's remaining test modules (gloo processes, nothing mocked):
tests/nn/tensor_parallel/{test_functional_,test_parallelizer}.py, tests/nn/expert_parallel/{test_experts,
test_hybrid_expert_parallel,test_expert_utils,test_expert_parallel_mapping}.py, tests/partitioning/test_profile.py,
tests/nn/data_parallel/test_data_parallel.py (parameter equality across replicas)."""
import copy

import pytest
import torch
import torch.distributed as dist
from torch import nn

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, ExpertParallel, TensorParallel
from pipegoose_b200.nn.expert_parallel import SwitchNoisePolicy, Top1Router
from pipegoose_b200.nn.expert_parallel.experts import Experts
from pipegoose_b200.nn.expert_parallel.parallel_mapping import ExpertParallelMapping
from pipegoose_b200.nn.expert_parallel.utils import get_num_local_experts
from pipegoose_b200.nn.tensor_parallel._functional import (
    broadcast_to_tensor_group,
    gather_to_tensor_group,
    reduce_to_tensor_group,
    scatter_to_tensor_group,
)
from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear
from pipegoose_b200.nn.tensor_parallel.parallelizer import (
    EmbeddingParallelizer,
    LayerNormParallelizer,
    LinearParallelizer,
    LMHeadParallelizer,
)
from pipegoose_b200.partitioning.profile import ProfileByMemory
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=99, hidden_size=32, n_layer=2, n_head=4)  # 99: not divisible by tp=2 -> padded vocab


# ------------------------------------------------------------------------------------------ TP autograd comm ops
def run_tp_functional(rank, world_size, port):
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    T = world_size
    # broadcast: identity forward, all-reduce backward
    x = torch.full((2, 4), float(rank + 1), requires_grad=True)
    y = broadcast_to_tensor_group(x, ctx)
    assert torch.equal(y, x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.full((2, 4), float(T)))
    # gather: all-gather forward along dim, local slice backward
    x = torch.full((2, 3), float(rank), requires_grad=True)
    y = gather_to_tensor_group(x, -1, ctx)
    assert y.shape == (2, 3 * T) and all(torch.equal(y[:, 3 * r:3 * r + 3], torch.full((2, 3), float(r))) for r in range(T))
    (y * torch.arange(3 * T, dtype=torch.float32)).sum().backward()
    assert torch.equal(x.grad, torch.arange(3 * T, dtype=torch.float32)[3 * rank:3 * rank + 3].expand(2, 3))
    # scatter: local slice forward, all-gather backward
    full = torch.arange(2 * 2 * T, dtype=torch.float32).view(2, 2 * T).requires_grad_(True)
    y = scatter_to_tensor_group(full, -1, ctx)
    assert torch.equal(y, full.detach()[:, 2 * rank:2 * rank + 2])
    (y * (rank + 1)).sum().backward()
    want = torch.cat([torch.full((2, 2), float(r + 1)) for r in range(T)], dim=-1)
    assert torch.equal(full.grad, want)
    # reduce: all-reduce forward, identity backward
    x = torch.full((3,), float(rank + 1), requires_grad=True)
    y = reduce_to_tensor_group(x, ctx)
    assert torch.equal(y, torch.full((3,), float(sum(range(1, T + 1)))))
    (2 * y).sum().backward()
    assert torch.equal(x.grad, torch.full((3,), 2.0))
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_tensor_parallel_autograd_comm_ops(tp):
    spawn(run_tp_functional, world_size=tp)


# ------------------------------------------------------------------------------------------ parallelizers
def run_parallelizers(rank, world_size, port, state):
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    T = world_size
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    full = copy.deepcopy(model)
    h, V = CFG["hidden_size"], CFG["vocab_size"]
    padded = (V + T - 1) // T * T

    name, emb = "transformer.word_embeddings", model.transformer.word_embeddings
    assert EmbeddingParallelizer.is_parallelizable(name, emb) and not LinearParallelizer.is_parallelizable(name, emb)
    EmbeddingParallelizer(name, emb, model, ctx).parallelize()
    assert isinstance(emb, ParallelEmbedding) and emb.weight.shape == (padded // T, h)
    assert (emb.vocab_start_idx, emb.vocab_end_idx) == (rank * padded // T, (rank + 1) * padded // T)
    rows = full.transformer.word_embeddings.weight[emb.vocab_start_idx:min(emb.vocab_end_idx, V)]
    assert torch.equal(emb.weight[:rows.shape[0]], rows) and emb.weight[rows.shape[0]:].abs().sum() == 0
    ids = torch.randint(0, V, (2, 5))
    assert torch.allclose(emb(ids), full.transformer.word_embeddings(ids), atol=1e-6)  # mask + all-reduce

    blk, fblk = model.transformer.h[0], full.transformer.h[0]
    qkv_name = "transformer.h.0.self_attention.query_key_value"
    assert LinearParallelizer.is_parallelizable(qkv_name, blk.self_attention.query_key_value)
    LinearParallelizer(qkv_name, blk.self_attention.query_key_value, model, ctx).parallelize()
    col = blk.self_attention.query_key_value
    assert isinstance(col, ColumnParallelLinear) and col.gather_output
    assert col.weight.shape == (3 * h // T, h) and col.bias.shape == (3 * h // T,)
    assert torch.equal(col.weight, fblk.self_attention.query_key_value.weight.chunk(T, 0)[rank])
    x = torch.randn(3, h)
    assert torch.allclose(col(x), fblk.self_attention.query_key_value(x), atol=1e-5)

    dense_name = "transformer.h.0.self_attention.dense"
    LinearParallelizer(dense_name, blk.self_attention.dense, model, ctx).parallelize()
    row = blk.self_attention.dense
    assert isinstance(row, RowParallelLinear) and row.weight.shape == (h, h // T) and row.bias.shape == (h,)
    assert torch.allclose(row(x), fblk.self_attention.dense(x), atol=1e-5)
    assert not LinearParallelizer.is_parallelizable("transformer.h.0.some_other_linear", nn.Linear(4, 4))

    ln_name = "transformer.h.0.input_layernorm"
    assert LayerNormParallelizer.is_parallelizable(ln_name, blk.input_layernorm)
    LayerNormParallelizer(ln_name, blk.input_layernorm, model, ctx).parallelize()
    assert isinstance(blk.input_layernorm, LayerNorm) and torch.equal(blk.input_layernorm.weight, fblk.input_layernorm.weight)
    assert torch.allclose(blk.input_layernorm(x), fblk.input_layernorm(x), atol=1e-6)

    # tied lm_head shares the (already sliced) embedding shard instead of being sliced again
    assert LMHeadParallelizer.is_parallelizable("lm_head", model.lm_head)
    LMHeadParallelizer("lm_head", model.lm_head, model, ctx).parallelize()
    assert isinstance(model.lm_head, ColumnParallelLinear) and model.lm_head.weight is emb.weight
    logits = model.lm_head(x)[..., :V]
    assert torch.allclose(logits, full.lm_head(x), atol=1e-5)
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2])
def test_parallelizers(tp):
    torch.manual_seed(0)
    spawn(run_parallelizers, world_size=tp, state=BloomForCausalLM(BloomConfig(**CFG)).state_dict())


# ------------------------------------------------------------------------------------------ experts
class _Expert(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.w = nn.Linear(d, d)

    def forward(self, x):
        return self.w(x)


def run_experts(rank, world_size, port, state, x, order, ref):
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)
    E = 4
    n_local = get_num_local_experts(E, ctx)
    assert n_local == E // world_size
    experts = Experts(n_local, _Expert(8), enable_tensor_parallel=False, parallel_context=ctx)
    assert all(getattr(p, "is_expert", False) for p in experts.parameters())
    for i, e in enumerate(experts.experts):  # distinct experts: global expert g multiplies by (g + 1)
        g = rank * n_local + i
        e.load_state_dict({"w.weight": state["w.weight"] * (g + 1), "w.bias": state["w.bias"]})
    xin = x.clone().requires_grad_(True)
    out = experts(xin, order)
    assert torch.allclose(out, ref, atol=1e-5)
    if out.requires_grad:  # a rank whose experts received no token at all contributes a constant zero
        out.sum().backward()
    # only experts that received tokens get gradients
    for i, e in enumerate(experts.experts):
        g = rank * n_local + i
        got_tokens = bool((order == g).any())
        assert (e.w.weight.grad is not None and e.w.weight.grad.abs().sum() > 0) == got_tokens
    ctx.destroy()


@pytest.mark.parametrize("tp", [1, 2, 4])
def test_experts_route_tokens_to_local_experts(tp):
    torch.manual_seed(0)
    base = _Expert(8)
    x = torch.randn(2, 6, 8)
    order = torch.tensor([0, 1, 2, 0, 1, 2, 0, 0, 1, 1, 2, 2])  # expert 3 receives nothing
    flat = x.view(-1, 8)
    ref = torch.stack([flat[i] @ (base.w.weight * (int(order[i]) + 1)).t() + base.w.bias for i in range(12)]).view(2, 6, 8)
    spawn(run_experts, world_size=tp, state=base.state_dict(), x=x, order=order, ref=ref.detach())


def test_expert_parallel_mapping_and_utils():
    assert ExpertParallelMapping.is_mlp("transformer.h.0.mlp")
    assert not ExpertParallelMapping.is_mlp("transformer.h.0.self_attention.dense")


def run_expert_data_parallel(rank, world_size, port, state, ids):
    """EP (over the TENSOR group) x DP: after backward, an expert's gradients are identical on the ranks that hold
    the SAME expert (the EXPERT_DATA group), dense gradients are identical on the DATA group."""
    ctx = init_parallel_context(rank, world_size, port, 2, 1, 2)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    torch.manual_seed(123)  # same router init everywhere
    router = Top1Router(SwitchNoisePolicy(), 4, 32)
    model = ExpertParallel(model, 4, mapping=[0], router=router, parallel_context=ctx).parallelize()
    model = TensorParallel(model, ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    model.eval()  # no routing noise: replicas differ only by their data
    local = ids.chunk(2)[ctx.get_local_rank(ParallelMode.DATA)]
    model(local, labels=local).loss.backward()
    reducer = getattr(model, "_pg_grad_reducer", None)
    n_expert = 0
    for name, p in model.named_parameters():
        g = p.main_grad if getattr(p, "main_grad", None) is not None else p.grad
        if g is None:
            continue
        mode = ParallelMode.EXPERT_DATA if getattr(p, "is_expert", False) else ParallelMode.DATA
        n_expert += int(getattr(p, "is_expert", False))
        group = ctx.get_group(mode)
        gathered = [torch.empty_like(g) for _ in range(ctx.get_world_size(mode))]
        dist.all_gather(gathered, g.contiguous(), group=group)
        for other in gathered[1:]:
            assert torch.allclose(gathered[0], other, atol=1e-6), name
    assert n_expert > 0 and reducer is not None
    ctx.destroy()


def test_expert_gradients_are_averaged_over_the_expert_data_group():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    spawn(run_expert_data_parallel, world_size=4, state=model.state_dict(), ids=torch.randint(0, 96, (4, 8)))


# ------------------------------------------------------------------------------------------ misc
def test_profile_by_memory():
    model = nn.Sequential(nn.Linear(16, 64), nn.ReLU(), nn.Linear(64, 8))
    sizes = ProfileByMemory(model).profile(torch.randn(4, 16))
    assert len(sizes) == 3 and all(isinstance(s, int) and s >= 0 for s in sizes)
    assert sizes[0] > sizes[1]  # a Linear with parameters outweighs the parameter-free ReLU


def run_dp_replicas(rank, world_size, port, state, ids):
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    model.load_state_dict(state)
    model = DataParallel(model, ctx).parallelize()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    for _ in range(2):
        local = ids.chunk(world_size)[rank]
        loss = model(local, labels=local).loss
        opt.zero_grad()
        loss.backward()  # the reducer exposes the averaged gradients as ``p.grad`` for the stock optimizer
        assert all(p.grad is not None for p in model.parameters())
        opt.step()
    for name, p in model.named_parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        assert torch.allclose(p.detach(), ref, atol=1e-7), name  # replicas stay bit-identical
    assert sum(int(not torch.equal(p.detach(), before[n])) for n, p in model.named_parameters()) > 10, "nothing was trained"
    ctx.destroy()


def test_data_parallel_replicas_stay_identical_with_a_stock_optimizer():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=96, hidden_size=32, n_layer=2, n_head=4))
    spawn(run_dp_replicas, world_size=2, state=model.state_dict(), ids=torch.randint(0, 96, (4, 8)))


@pytest.mark.parametrize("D", [80, 40, 96])
def test_padded_head_attention_is_exact(D):
    """Head sizes without a native tile shape (bloom-3b: D=80) run on the D=64/128 flash kernel after zero-padding
    every head and passing the softmax scale of the true width.  The padding algebra is checked here against the
    unpadded PyTorch reference, forward and backward (the kernel itself is covered by the GPU tests at D=64/128)."""
    import math

    from pipegoose_b200.ops.attention import alibi_attention_reference, pad_heads, padded_head_dim, unpad_heads
    from pipegoose_b200.ops.kernels import alibi_slopes

    B, S, H = 2, 16, 4
    DP = padded_head_dim(D)
    assert DP in (64, 128) and DP >= D
    torch.manual_seed(0)
    qkv = torch.randn(B * S, H * 3 * D, dtype=torch.float64, requires_grad=True)
    slopes = alibi_slopes(H)
    ref = alibi_attention_reference(qkv, slopes, B, S, H, D)
    qkv2 = qkv.detach().clone().requires_grad_(True)
    out_p = alibi_attention_reference(pad_heads(qkv2, H, 3, D, DP), slopes, B, S, H, DP, softmax_scale=1.0 / math.sqrt(D))
    got = unpad_heads(out_p, H, 1, D, DP)
    assert torch.allclose(got, ref, atol=1e-6)
    # the padded output columns are exactly zero (v was padded with zeros)
    assert out_p.view(B * S, H, DP)[..., D:].abs().max() == 0
    g = torch.randn_like(ref)
    ref.backward(g)
    got.backward(g)
    assert torch.allclose(qkv2.grad, qkv.grad, atol=1e-6)


def test_remaining_reference_symbols():
    """Smaller public names of the reference that the other tests do not touch."""
    import io

    from pipegoose_b200.nn.pipeline_parallel._worker import BaseWorkerManager, WorkerManager
    from pipegoose_b200.nn.pipeline_parallel.exception import PipelineError, PipelineInputNotRequiresGrad, PipelineNoSavedInput
    from pipegoose_b200.nn.pipeline_parallel.partitioner import BasePartitioner, PartitionPolicy, UniformPartitioner, _get_partitioner
    from pipegoose_b200.nn.pipeline_parallel.scheduler import GPipeScheduler
    from pipegoose_b200.nn.pipeline_parallel.sync.progress_tracker import get_progresses_from_pipeline_context
    from pipegoose_b200.partitioning.profile import ProfileByMemory, ProfileStrategy
    from pipegoose_b200.trainer import DistributedLogger

    assert _get_partitioner(PartitionPolicy.UNIFORM) is UniformPartitioner and issubclass(UniformPartitioner, BasePartitioner)
    assert issubclass(WorkerManager, BaseWorkerManager) and issubclass(ProfileByMemory, ProfileStrategy)
    assert issubclass(PipelineInputNotRequiresGrad, PipelineError)
    assert "micro-batch 3, partition 1" in str(PipelineNoSavedInput(3, 1))

    class Ctx:  # only ``schedules`` is read
        schedules = GPipeScheduler(2, 2).get_schedules()

    table = get_progresses_from_pipeline_context(Ctx)
    assert len(table) == len(Ctx.schedules) and table[0] == {(0, 0): False}
    assert all(v is False for clock in table.values() for v in clock.values())

    stream = io.StringIO()
    log = DistributedLogger(stream=stream).set_level("WARNING")
    log.info("quiet"), log.warning("loud"), log.log("also loud", "ERROR")
    assert "quiet" not in stream.getvalue() and "loud" in stream.getvalue() and "also loud" in stream.getvalue()
    assert [lvl for lvl, _ in log.records] == ["INFO", "WARNING", "ERROR"]
