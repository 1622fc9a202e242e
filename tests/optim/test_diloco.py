"""DiLoCo outer optimizer (optim/diloco.py): workers train without communication for H steps, then apply the averaged
displacement with Nesterov momentum — compared with a sequential simulation of the same algorithm."""
import copy

import pytest
import torch

from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import TensorParallel
from pipegoose_b200.optim import FusedAdam, clip_grad_norm_
from pipegoose_b200.optim.diloco import DiLoCoOptimizer
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)
H, ROUNDS, WORKERS = 2, 3, 2
OUTER_LR, MU = 0.7, 0.9


def _data(worker, step):
    g = torch.Generator().manual_seed(1000 * worker + step)
    return torch.randint(0, 96, (2, 8), generator=g)


def _simulate(state, fused):
    """The algorithm, written out for WORKERS replicas in one process."""
    models = []
    for _ in range(WORKERS):
        m = BloomForCausalLM(BloomConfig(**CFG))
        m.load_state_dict(state)
        models.append(m)
    opts = [FusedAdam(m.parameters(), lr=1e-2) if fused else torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.0)
            for m in models]
    names = [n for n, _ in models[0].named_parameters()]
    anchor = {n: p.detach().clone() for n, p in models[0].named_parameters()}
    vel = {n: torch.zeros_like(v) for n, v in anchor.items()}
    for r in range(ROUNDS):
        for k, (m, o) in enumerate(zip(models, opts)):
            for h in range(H):
                ids = _data(k, r * H + h)
                loss = m(ids, labels=ids).loss
                o.zero_grad()
                loss.backward()
                o.step()
        with torch.no_grad():
            for n in names:
                delta = torch.stack([anchor[n] - dict(m.named_parameters())[n] for m in models]).mean(0)
                vel[n] = MU * vel[n] + delta
                anchor[n] = anchor[n] - OUTER_LR * (delta + MU * vel[n])
            for m in models:
                for n, p in m.named_parameters():
                    p.copy_(anchor[n])
            for o in opts:  # FusedAdam keeps fp32 masters: continue from the shared parameters
                if isinstance(o, FusedAdam):
                    o.master.copy_(o.flat.flat_param)
    return {n: v.clone() for n, v in anchor.items()}


def run_diloco(rank, world_size, port, tp, fused, state, want):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, WORKERS)
    worker = ctx.get_local_rank(ParallelMode.DATA)
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    names = {id(p): n for n, p in model.named_parameters()}
    model = TensorParallel(model, ctx).parallelize()          # no DataParallel: workers do not share gradients
    inner = FusedAdam(model.parameters(), lr=1e-2) if fused else torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0)
    optim = DiLoCoOptimizer(inner, ctx, inner_steps=H, outer_lr=OUTER_LR, outer_momentum=MU)
    for step in range(H * ROUNDS):
        ids = _data(worker, step)
        loss = model(ids, labels=ids).loss
        optim.zero_grad()
        loss.backward()
        norm = clip_grad_norm_(optim, 1e9, ctx)     # (unwraps to the worker's inner optimizer; 1e9: never bites)
        assert torch.isfinite(norm) and norm > 0
        optim.step()
    assert optim.outer_step_count == ROUNDS and optim.local_step == H * ROUNDS
    for p in model.parameters():
        n = names.get(id(p))
        if n is not None and p.shape == want[n].shape:          # unsharded parameters compare directly
            assert torch.allclose(p.detach(), want[n], atol=2e-5), n
    sd = optim.state_dict()
    optim.load_state_dict(copy.deepcopy(sd))
    assert optim.outer_step_count == ROUNDS
    ctx.destroy()


@pytest.mark.parametrize("tp,fused", [(1, False), (1, True), (2, True)])
def test_diloco_matches_sequential_simulation(tp, fused):
    torch.manual_seed(0)
    state = copy.deepcopy(BloomForCausalLM(BloomConfig(**CFG)).state_dict())
    want = _simulate(state, fused)
    spawn(run_diloco, world_size=tp * WORKERS, tp=tp, fused=fused, state=state, want=want)
