"""The gradient reducer / ZeRO-1 optimizer logic that is only active with the NVLink engine — a head bucket for the small
parameters, multi-bucket tail launches, ZeRO slices and parameter all-gather regions cut around the head — exercised on
CPU with a stand-in engine that speaks the engine's interface over gloo."""
import copy

import pytest
import torch
import torch.distributed as dist

from pipegoose_b200.core.flat_state import FlatModelState
from pipegoose_b200.core.grad_reducer import GradReducer
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import DataParallel
from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
from pipegoose_b200.testing.utils import init_parallel_context, spawn

CFG = dict(vocab_size=96, hidden_size=32, n_layer=2, n_head=4)


class _Done:
    def wait(self):
        return True


class StandInEngine:
    """What ``ops.comm.FusedDPEngine`` offers the reducer and the optimizer, on CPU tensors and gloo collectives."""

    inline_start = None

    def __init__(self, ctx):
        self.group = ctx.get_group(ParallelMode.DATA)
        self.world = ctx.get_world_size(ParallelMode.DATA)
        self.rank = ctx.get_local_rank(ParallelMode.DATA)
        self.calls = []

    def allocate(self, numel, param_dtype, grad_dtype):
        return torch.zeros(numel, dtype=param_dtype), torch.zeros(numel, dtype=grad_dtype)

    def begin_overlap(self):
        pass

    def end_overlap(self):
        pass

    def enable_inline(self, start):
        return False

    def reduce_bucket(self, view, mode, tail=False, bucket_numel=0):
        n = view.numel()
        step = bucket_numel or n
        self.calls.append({"numel": n, "tail": tail, "bucket_numel": bucket_numel})
        for b0 in range(0, n, step):
            sub = view[b0:b0 + step]
            seg = sub.numel() // self.world
            full = sub.clone()
            dist.all_reduce(full, group=self.group)
            if mode == "reduce_scatter":   # only this rank's slice of every bucket holds the average afterwards
                sub[self.rank * seg:(self.rank + 1) * seg] = full[self.rank * seg:(self.rank + 1) * seg] / self.world
            else:
                sub.copy_(full / self.world)
        return _Done()

    def all_gather_params(self, flat_param, bucket_numel, head=0):
        n = flat_param.numel()
        regions = ([(0, head)] if head else []) + [(s, min(n, s + bucket_numel)) for s in range(head, n, bucket_numel)]
        for s, e in regions:
            seg = (e - s) // self.world
            mine = flat_param[s + self.rank * seg:s + (self.rank + 1) * seg].clone()
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            flat_param[s:e] = torch.cat(parts)


def run_engine(rank, world_size, port, state, ids, ref_state, cfg=None, bucket_mb=0.02):
    cfg = cfg or CFG
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size)
    engine = StandInEngine(ctx)

    def make_flat_state(self):
        st = FlatModelState(self.module.parameters(), pad_to_multiple_of=self.dp, buffer_factory=engine.allocate)
        self.module._flat_state = st
        self._fused = engine
        return st

    GradReducer._make_flat_state = make_flat_state
    model = BloomForCausalLM(BloomConfig(**cfg))
    model.load_state_dict(state)
    model = DataParallel(model, ctx, bucket_size_mb=bucket_mb).parallelize()
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-3), ctx)
    local = ids.chunk(world_size)[rank]
    for _ in range(3):
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
    reducer = model._pg_grad_reducer
    flat = reducer.flat
    assert 0 < reducer.head == flat.matrix_start < flat.numel                       # small parameters got their own bucket
    assert reducer.buckets[0].end == reducer.head and len(reducer.buckets) >= 3
    assert all(p.dim() < 2 for b in reducer.buckets[:1] for p in b.params)
    segs = optim.optim._segments
    assert segs[0][1] <= reducer.head and segs[1][0] >= reducer.head                # ZeRO slices are cut around the head
    got = model.state_dict()
    for k, v in ref_state.items():
        assert torch.allclose(got[k], v, atol=2e-4), k
    # opt-in merging of runs of unlaunched buckets into one multi-bucket launch (bucket-wise slices are preserved)
    with model.no_sync():
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()                      # nothing was reduced
    want = flat.flat_grad.clone()
    dist.all_reduce(want, group=ctx.get_group(ParallelMode.DATA))
    want /= world_size
    reducer.MERGE_TAIL = True
    engine.calls.clear()
    reducer._launch_tail()
    merged = [c for c in engine.calls if c["bucket_numel"] and c["numel"] > c["bucket_numel"]]
    assert merged and all(c["tail"] for c in merged) and len(engine.calls) < len(reducer.buckets)
    for s0, e0 in optim.optim._segments:     # this rank's slice of every bucket holds the average
        assert torch.allclose(flat.flat_grad[s0:e0], want[s0:e0], atol=1e-6)
    ctx.destroy()


@pytest.mark.parametrize("world,cfg,bucket_mb", [
    (2, CFG, 0.02),
    (3, dict(vocab_size=100, hidden_size=48, n_layer=2, n_head=4), 0.01),     # sizes that do not divide by the group size
    (4, dict(vocab_size=50, hidden_size=32, n_layer=1, n_head=2), 0.008),     # buckets smaller than the embedding table
])
def test_head_bucket_merged_tail_and_zero_slices_with_an_engine(world, cfg, bucket_mb):
    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(**cfg))
    state = copy.deepcopy(model.state_dict())
    ids = torch.randint(0, cfg["vocab_size"], (12, 16))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for _ in range(3):
        loss = model(ids, labels=ids).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
    spawn(run_engine, world_size=world, state=state, ids=ids, ref_state={k: v.clone() for k, v in model.state_dict().items()},
          cfg=cfg, bucket_mb=bucket_mb)
