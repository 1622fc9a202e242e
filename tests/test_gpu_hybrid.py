"""Hybrid parallelism on real GPUs (NCCL + the fused NVLink kernels): TP x DP (+ZeRO-1), PP (1F1B over NCCL
p2p) and TP x PP against the single-GPU bf16 model — the compositions BASELINE.json's configs are made of.
Loss trajectories are compared over three optimizer steps (bf16 tolerances)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(vocab_size=4096, hidden_size=256, n_layer=4, n_head=4)
SEQ = 256


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _reference(ids, n_microbatches, dp, steps=3):
    """Single-GPU trajectory with the loss definition of the parallel run: mean over data-parallel replicas of
    the mean over micro-batches of the per-micro-batch mean loss."""
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.optim import FusedAdam

    torch.manual_seed(0)
    ref = BloomForCausalLM(BloomConfig(**CFG))
    state = copy.deepcopy(ref.state_dict())
    model = ref.to(torch.bfloat16).cuda()
    opt = FusedAdam(model.parameters(), lr=1e-3)
    losses = []
    chunks = [mb for rep in ids.cuda().chunk(dp) for mb in rep.chunk(n_microbatches)]
    for _ in range(steps):
        opt.zero_grad()
        total = 0.0
        for mb in chunks:
            loss = model(mb, labels=mb).loss / len(chunks)
            loss.backward()
            total += loss.item()
        opt.step()
        losses.append(total)
    return state, losses


def run_hybrid(rank, world_size, port, tp, pp, dp, n_microbatches, state, ids, ref_losses):
    import torch.distributed as dist

    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel, PipelineParallel, TensorParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, tp, pp, dp, backend="nccl")
    model = BloomForCausalLM(BloomConfig(**CFG))
    model.load_state_dict(state)
    model = model.to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=n_microbatches, parallel_context=ctx).parallelize()
    model = DataParallel(model, ctx, bucket_size_mb=1.0).parallelize()
    model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-3), ctx)
    local = ids.chunk(dp)[ctx.get_local_rank(ParallelMode.DATA)].cuda()
    losses = []
    for _ in range(len(ref_losses)):
        out = model(local, labels=local)
        optim.zero_grad()
        out.loss.backward()
        optim.step()
        losses.append(float(out.loss.item()) if out.loss is not None else 0.0)
    # the loss lives on the last pipeline stage; average it over the data-parallel replicas
    t = torch.tensor(losses, device="cuda")
    if not ctx.is_last_rank(ParallelMode.PIPELINE):
        t.zero_()
    dist.all_reduce(t)
    n_holders = world_size // pp  # ranks of the last stage (every TP rank reports the same loss)
    mean_losses = (t / n_holders).tolist()
    for a, b in zip(mean_losses, ref_losses):
        assert abs(a - b) < 6e-2, (mean_losses, ref_losses)
    assert mean_losses[-1] < mean_losses[0], mean_losses
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,dp,mb", [(2, 1, 2, 1), (1, 2, 1, 2), (2, 2, 1, 2), (1, 2, 2, 2)])
def test_hybrid_matches_single_gpu(tp, pp, dp, mb):
    _need_gpus(tp * pp * dp)
    from pipegoose_b200.testing.utils import spawn

    torch.manual_seed(1)
    ids = torch.randint(0, CFG["vocab_size"], (4 * dp, SEQ))
    state, ref_losses = _reference(ids, mb, dp)
    spawn(run_hybrid, world_size=tp * pp * dp, tp=tp, pp=pp, dp=dp, n_microbatches=mb, state=state, ids=ids,
          ref_losses=ref_losses)
