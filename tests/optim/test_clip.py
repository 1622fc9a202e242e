"""Global gradient-norm clipping (optim/clip.py): the norm computed under TP / PP / DP / ZeRO-1 layouts equals the norm
of the single-process model's gradient, and a clipped step lands on the same parameters."""
import copy

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam, clip_grad_norm_
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=4, n_head=4)
MAX_NORM = 0.05   # far below the actual norm: the clip always bites
LR = 1e-2
EPS = 1e-3    # a large Adam epsilon makes the update sensitive to the gradient scale (plain Adam is nearly invariant)


def _reference(state, ids, chunks, fused):
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    opt = FusedAdam(model.parameters(), lr=LR, eps=EPS) if fused else torch.optim.SGD(model.parameters(), lr=1.0)
    opt.zero_grad()
    for mb in chunks:
        (model(mb, labels=mb).loss / len(chunks)).backward()
    if fused:
        norm = clip_grad_norm_(opt, MAX_NORM, _SingleRank())
    else:
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), MAX_NORM)
    opt.step()
    return float(norm), {k: v.detach().clone() for k, v in model.state_dict().items()}


class _SingleRank:
    """Stands in for a ParallelContext in the single-process reference."""

    def get_world_size(self, mode):
        return 1


def run_clip(rank, world_size, port, tp, pp, dp, fused, n_mb, state, ids, ref_norm, ref_state):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = TensorParallel(model, ctx).parallelize()
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=n_mb, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx).parallelize()
    inner = FusedAdam(model.parameters(), lr=LR, eps=EPS) if fused else torch.optim.SGD(model.parameters(), lr=1.0)
    optim = DistributedOptimizer(inner, ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    out = model(local, labels=local)
    optim.zero_grad()
    out.loss.backward()
    norm = optim.clip_grad_norm_(MAX_NORM)
    assert abs(float(norm) - ref_norm) < 2e-4 * max(1.0, ref_norm), (float(norm), ref_norm)
    optim.step()
    # unsharded parameters can be compared directly with the single-process model after the clipped step
    for p in model.parameters():
        n = names.get(id(p))
        if n is not None and p.numel() > 0 and p.shape == ref_state[n].shape:
            assert torch.allclose(p.detach().float(), ref_state[n], atol=3e-5), n
    # and the sharded ones through the loss of the next forward
    loss2 = model(local, labels=local).loss.detach().float().reshape(1).clone()
    dist.all_reduce(loss2)
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,dp,fused", [(2, 1, 2, True), (1, 1, 2, False), (2, 2, 1, True), (1, 2, 2, True)])
def test_clipped_step_matches_single_process(tp, pp, dp, fused):
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    ids = torch.randint(0, 96, (8, 8))
    n_mb = 2
    chunks = [mb for rep in ids.chunk(dp) for mb in (rep.chunk(n_mb) if pp > 1 else [rep])]
    ref_norm, ref_state = _reference(state, ids, chunks, fused)
    assert ref_norm > 10 * MAX_NORM
    spawn(run_clip, world_size=tp * pp * dp, tp=tp, pp=pp, dp=dp, fused=fused, n_mb=n_mb, state=state, ids=ids,
          ref_norm=ref_norm, ref_state=ref_state)


def test_single_process_norm_matches_torch():
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**CFG))
    ids = torch.randint(0, 96, (4, 8))
    model(ids, labels=ids).loss.backward()
    want = torch.nn.utils.clip_grad_norm_([p for p in model.parameters()], 1e9)
    got = clip_grad_norm_(model, 1e9, _SingleRank())
    assert torch.allclose(got, want, rtol=1e-5)


# ------------------------------------------------------------------ MoE: gradient norm across layouts
from pipegoose_b200.nn import ExpertParallel  # noqa: E402
from pipegoose_b200.nn.expert_parallel import ExpertLoss, SwitchNoisePolicy, Top1Router  # noqa: E402


def run_moe_norm(rank, world_size, port, tp, dp, state, gate_state, ids, out_file):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, dp)
    model = BloomForCausalLM(BloomConfig(**CFG)); model.load_state_dict(state)
    router = Top1Router(SwitchNoisePolicy(), 4, 32); router.load_state_dict(gate_state)
    model = ExpertParallel(model, 4, mapping=[1], router=router, parallel_context=ctx).parallelize()
    layer = model.transformer.h[1].mlp
    first = ctx.get_local_rank(ParallelMode.TENSOR) * len(layer.experts)
    for i, e in enumerate(layer.experts):
        g = torch.Generator().manual_seed(500 + first + i)
        for p in e.parameters():
            p.data = p.data + 0.05 * torch.randn(p.shape, generator=g)
    model = TensorParallel(model, ctx).parallelize(); model = DataParallel(model, ctx).parallelize()
    model.eval()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-2), ctx)
    loss_fn = ExpertLoss(lambda out: out.loss, aux_weight=0.01, z_weight=0.001)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)]
    loss = loss_fn(model(local, labels=local)); optim.zero_grad(); loss.backward()
    norm = optim.clip_grad_norm_(1e9)
    if rank == 0: torch.save(norm, out_file)
    ctx.destroy()
def test_global_gradient_norm_is_layout_independent_with_sharded_experts(tmp_path):
    """Distinct experts sharded over the tensor group, router losses included, ZeRO-1 slices: the same norm in every layout."""
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    gate_state = copy.deepcopy(Top1Router(SwitchNoisePolicy(), 4, 32).state_dict())
    ids = torch.randint(0, 96, (4, 8)); res = {}
    for name, (tp, dp) in {"a": (1, 1), "b": (2, 1), "c": (2, 2), "d": (1, 2)}.items():
        f = str(tmp_path / name); spawn(run_moe_norm, world_size=tp * dp, tp=tp, dp=dp, state=state, gate_state=gate_state, ids=ids, out_file=f)
        res[name] = torch.load(f).item()
    assert abs(res["a"] - res["b"]) < 1e-4 * res["a"], res
    assert abs(res["c"] - res["d"]) < 1e-4 * res["d"], res
