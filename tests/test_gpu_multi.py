"""Multi-GPU numerics (skipped on single-GPU boxes): fused all-gather->GEMM / GEMM->reduce-scatter
kernels against NCCL + torch.matmul, and the TP2 sequence-parallel Bloom against the single-GPU model."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def run_fused_tp(rank, world_size, port):
    import torch.distributed as dist

    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.parallel.tp_comm import TensorParallelComm
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1, backend="nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    comm = TensorParallelComm(ctx, fused=True)
    comm.enable_fused()
    assert comm.fused
    T = world_size
    torch.manual_seed(0)  # same full tensors on every rank
    M, K, N = 1024 * T, 512, 768
    x_full = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
    x_shard = x_full.chunk(T)[rank].contiguous()

    def rel(a, b):
        return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-6)

    for it in range(3):  # several epochs: exercises the double-buffered flags/counters
        # ---- all-gather -> GEMM (column-parallel forward), with bias and with GELU + pre-activation
        y, gathered = comm.ag_gemm(x_shard, w, bias)
        assert torch.equal(gathered, x_full)
        assert rel(y, x_full.float() @ w.float().t() + bias.float()) < 2e-2
        holder = {}
        y2, _ = comm.ag_gemm(x_shard, w, bias, gelu=True, aux_holder=holder)
        pre = x_full.float() @ w.float().t() + bias.float()
        from pipegoose_b200.ops import kernels as Kk

        assert rel(holder["aux"], pre) < 2e-2 and rel(y2, Kk.gelu_tanh(pre)) < 2e-2
        # ---- GEMM -> reduce-scatter (row-parallel forward) with bias + residual
        a_full = torch.randn(M, K * T, device=dev, dtype=torch.bfloat16)
        w_row = torch.randn(N, K * T, device=dev, dtype=torch.bfloat16) * 0.05
        res = torch.randn(M // T, N, device=dev, dtype=torch.bfloat16)
        a_loc = a_full[:, rank * K:(rank + 1) * K].contiguous()
        w_loc = w_row[:, rank * K:(rank + 1) * K].contiguous()
        out = comm.gemm_rs(a_loc, w_loc, bias, res)
        want = (a_full.float() @ w_row.float().t()).chunk(T)[rank] + bias.float() + res.float()
        assert rel(out, want) < 2e-2
        # ---- dgrad forms
        dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        dx = comm.gemm_rs_nn(dy, w)  # partial [M, K] summed over ranks then scattered
        full = dy.float() @ w.float()
        assert rel(dx, (full * T).chunk(T)[rank]) < 2e-2  # every rank holds the same dy/w here -> T x
        dy_shard = dy.chunk(T)[rank].contiguous()
        da, dy_g = comm.ag_gemm_nn(dy_shard, w)
        assert torch.equal(dy_g, dy) and rel(da, full) < 2e-2
    torch.cuda.synchronize()
    dist.barrier()
    ctx.destroy()


@pytest.mark.parametrize("world", [2, 4])
def test_fused_tp_kernels(world):
    _need_gpus(world)
    from pipegoose_b200.testing.utils import spawn

    spawn(run_fused_tp, world_size=world)


def run_tp_bloom(rank, world_size, port, fused, state, ids, ref_loss, ref_gnorm):
    import torch.distributed as dist

    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import TensorParallel
    from pipegoose_b200.testing.utils import init_parallel_context

    os.environ["PIPEGOOSE_B200_FUSED_TP"] = "1" if fused else "0"
    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1, backend="nccl")
    cfg = BloomConfig(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    model.load_state_dict(state)
    model = model.to(torch.bfloat16)
    model = TensorParallel(model, ctx).parallelize()
    model.to("cuda")
    ids = ids.cuda()
    loss = model(ids, labels=ids).loss
    loss.backward()
    assert abs(loss.item() - ref_loss) < 3e-2, (loss.item(), ref_loss)
    # global gradient norm (sharded params: sum of squares over ranks; replicated params: their partial gradients
    # were already summed over the tensor group by TensorParallel's TensorPartialGradSync, count them once)
    sq = torch.zeros((), device="cuda")
    for n, p in model.named_parameters():
        g = p.grad.float()
        if getattr(p, "tp_partial_grad", False):
            sq += g.pow(2).sum() / world_size
        elif hasattr(p, "parallel_metadata"):
            sq += g.pow(2).sum()
        else:
            sq += g.pow(2).sum() / world_size
    dist.all_reduce(sq)
    assert abs(sq.sqrt().item() - ref_gnorm) / ref_gnorm < 5e-2, (sq.sqrt().item(), ref_gnorm)
    ctx.destroy()


@pytest.mark.parametrize("fused", [False, True])
def test_tp2_bloom_matches_single_gpu(fused):
    _need_gpus(2)
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.testing.utils import spawn

    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4)
    ref = BloomForCausalLM(cfg)
    state = copy.deepcopy(ref.state_dict())
    ids = torch.randint(0, 4096, (2, 256))
    model = copy.deepcopy(ref).to(torch.bfloat16).cuda()
    loss = model(ids.cuda(), labels=ids.cuda()).loss
    loss.backward()
    gnorm = torch.sqrt(sum(p.grad.float().pow(2).sum() for p in model.parameters())).item()
    spawn(run_tp_bloom, world_size=2, fused=fused, state=state, ids=ids, ref_loss=loss.item(), ref_gnorm=gnorm)


def run_fused_moe(rank, world_size, port, top_k):
    import torch.distributed as dist
    from torch import nn

    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.models.bloom import BloomConfig, BloomMLP
    from pipegoose_b200.nn.expert_parallel import ExpertContext, Top1Router, Top2Router
    from pipegoose_b200.ops import kernels as Kk
    from pipegoose_b200.ops.moe import FusedExpertLayer
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1, backend="nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    T, E, h, n = world_size, 4, 256, 512
    torch.manual_seed(0)
    expert = BloomMLP(BloomConfig(hidden_size=h, n_head=4))
    router = (Top1Router if top_k == 1 else Top2Router)(None, E, h, expert_capacity=(8.0, 8.0))  # no drops
    layer = FusedExpertLayer(E, expert, router, ctx).to(torch.bfloat16).to(dev)
    El = E // T
    # distinct experts, identical construction on every rank
    g = torch.Generator().manual_seed(5)
    W1 = (torch.randn(E, 4 * h, h, generator=g) * 0.05).to(torch.bfloat16)
    B1 = (torch.randn(E, 4 * h, generator=g) * 0.05).to(torch.bfloat16)
    W2 = (torch.randn(E, h, 4 * h, generator=g) * 0.05).to(torch.bfloat16)
    B2 = (torch.randn(E, h, generator=g) * 0.05).to(torch.bfloat16)
    Wg = (torch.randn(E, h, generator=g) * 0.5).to(torch.bfloat16)
    with torch.no_grad():
        sl = slice(rank * El, (rank + 1) * El)
        layer.w1.copy_(W1[sl]); layer.b1.copy_(B1[sl]); layer.w2.copy_(W2[sl]); layer.b2.copy_(B2[sl])
        layer.router.gate.weight.copy_(Wg); layer.router.gate.bias.zero_()
    X = torch.randn(T * n, h, generator=g).to(torch.bfloat16)
    R = torch.randn(T * n, h, generator=g).to(torch.bfloat16)
    DY = torch.randn(T * n, h, generator=g).to(torch.bfloat16)
    x = X[rank * n:(rank + 1) * n].to(dev).requires_grad_(True)
    res = R[rank * n:(rank + 1) * n].to(dev).requires_grad_(True)
    layer.eval()
    for _ in range(2):  # two calls: exercises buffer parity / reset
        ExpertContext.get_instance().pop_all_aux_loss(); ExpertContext.get_instance().pop_all_z_loss()
        y = layer(x, res)
    y.backward(DY[rank * n:(rank + 1) * n].to(dev))
    torch.cuda.synchronize()
    # ---- dense fp32 reference over all tokens / all experts
    Xf = X.float().to(dev).requires_grad_(True)
    W1f, B1f, W2f, B2f, Wgf = (t.float().to(dev).requires_grad_(True) for t in (W1, B1, W2, B2, Wg))
    probs = torch.softmax(Xf @ Wgf.t(), dim=-1)
    tp, ti = torch.topk(probs, top_k, dim=-1)
    out = R.float().to(dev).clone()
    for e in range(E):
        he = Kk.gelu_tanh(Xf @ W1f[e].t() + B1f[e]) @ W2f[e].t() + B2f[e]
        w = (tp * (ti == e)).sum(-1, keepdim=True)
        out = out + w * he
    out.backward(DY.float().to(dev))

    def rel(a, b):
        return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-6)

    sl_t = slice(rank * n, (rank + 1) * n)
    assert rel(y, out[sl_t]) < 3e-2, rel(y, out[sl_t])
    assert rel(x.grad, Xf.grad[sl_t]) < 6e-2, rel(x.grad, Xf.grad[sl_t])
    assert rel(res.grad, DY[sl_t].to(dev)) < 1e-3
    sl = slice(rank * El, (rank + 1) * El)
    assert rel(layer.w1.grad, W1f.grad[sl]) < 6e-2 and rel(layer.w2.grad, W2f.grad[sl]) < 6e-2
    assert rel(layer.b1.grad, B1f.grad[sl]) < 6e-2 and rel(layer.b2.grad, B2f.grad[sl]) < 6e-2
    gg = layer.router.gate.weight.grad.float().clone()
    dist.all_reduce(gg)
    assert rel(gg, Wgf.grad) < 1e-1, rel(gg, Wgf.grad)
    ctx.destroy()


@pytest.mark.parametrize("top_k", [1, 2])
def test_fused_moe_layer(top_k):
    _need_gpus(2)
    from pipegoose_b200.testing.utils import spawn

    spawn(run_fused_moe, world_size=2, top_k=top_k)


def run_dp_zero(rank, world_size, port, fused_dp, state, ids, ref_losses):
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.testing.utils import init_parallel_context

    os.environ["PIPEGOOSE_B200_FUSED_DP"] = "1" if fused_dp else "0"
    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size, backend="nccl")
    cfg = BloomConfig(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4)
    model = BloomForCausalLM(cfg)
    model.load_state_dict(state)
    model = model.to(torch.bfloat16)
    model = DataParallel(model, ctx, bucket_size_mb=1.0).parallelize()
    model.to("cuda")
    optim = DistributedOptimizer(FusedAdam(model.parameters(), lr=1e-3), ctx)
    local = ids.chunk(world_size)[rank].cuda()
    losses = []
    for _ in range(len(ref_losses)):
        loss = model(local, labels=local).loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        losses.append(loss.item())
    assert (model._pg_grad_reducer._fused is not None) == fused_dp
    import torch.distributed as dist

    t = torch.tensor(losses, device="cuda")
    dist.all_reduce(t)
    mean_losses = (t / world_size).tolist()
    for a, b in zip(mean_losses, ref_losses):
        assert abs(a - b) < 5e-2, (mean_losses, ref_losses)
    # replicas hold identical parameters after ZeRO-1's all-gather
    flat = model._flat_state.flat_param.float()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref)
    ctx.destroy()


@pytest.mark.parametrize("fused_dp,world", [(False, 2), (True, 2), (True, 4)])
def test_dp_zero1_matches_single_gpu(fused_dp, world):
    _need_gpus(world)
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.testing.utils import spawn

    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4)
    ref = BloomForCausalLM(cfg)
    state = copy.deepcopy(ref.state_dict())
    ids = torch.randint(0, 4096, (2 * world, 256))
    model = copy.deepcopy(ref).to(torch.bfloat16).cuda()
    opt = FusedAdam(model.parameters(), lr=1e-3)
    ref_losses = []
    for _ in range(3):
        loss = model(ids.cuda(), labels=ids.cuda()).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    spawn(run_dp_zero, world_size=world, fused_dp=fused_dp, state=state, ids=ids, ref_losses=ref_losses)
