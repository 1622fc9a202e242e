"""Multi-GPU tests (skipped on single-GPU boxes) of the NVLink engine pieces added in round 2:

* the symmetric workspace on the VMM API with an NVSwitch multicast object (``multimem.ld_reduce`` / ``multimem.st``),
* the in-kernel data-parallel gradient reduce-scatter (wgrad GEMM epilogue, embedding backward, gradient folds adding
  straight into the owner rank's ZeRO-1 slice) against plain PyTorch fp32 sums,
* a ZeRO-1 training run with the in-kernel reduce-scatter against the bucketed reducer and against one GPU.
"""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


# ------------------------------------------------------------------------------------------------ workspace / NVLS
def run_workspace(rank, world_size, port):
    import torch.distributed as dist

    from pipegoose_b200.distributed import symmetric as S
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.ops import native
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1, backend="nccl")
    n = 1 << 16
    ws = S.SymmetricWorkspace(ctx, ParallelMode.TENSOR, 3 * n * 4)
    print(f"[rank {rank}] symmetric workspace backend: {ws.backend}, mc_ptr {hex(ws.mc_ptr)}", flush=True)
    a = ws.local_tensor(0, (n,), torch.float32)
    b = ws.local_tensor(n * 4, (n,), torch.float32)
    a.copy_(torch.arange(n, device="cuda", dtype=torch.float32) * (rank + 1))
    b.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    # peer mappings: every rank pushes its quarter of region c (bf16 view) into all peers with the all-gather kernel
    c = ws.local_tensor(2 * n * 4, (n,), torch.float32)
    c.zero_()
    seg = n // world_size
    c[rank * seg:(rank + 1) * seg] = float(rank + 1)
    torch.cuda.synchronize()
    dist.barrier()
    native().allgather_bf16([ws.data_ptr(p, 2 * n * 4) for p in range(world_size)], rank, 2 * n, 2 * n,
                            [ws.sig_ptr(p, S.SIG_BARRIER) for p in range(world_size)], 1, 0, 0)
    torch.cuda.synchronize()
    want_c = torch.cat([torch.full((seg,), float(r + 1), device="cuda") for r in range(world_size)])
    assert torch.equal(c, want_c), "peer stores did not arrive"
    if ws.mc_ptr:
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        # multimem.ld_reduce over the replicas of `a`; rank 0 multicasts the result into every replica of `b`
        native().multimem_selftest(ws.mc_data_ptr(0), ws.mc_data_ptr(n * 4) if rank == 0 else 0, out)
        torch.cuda.synchronize()
        dist.barrier()
        want = torch.arange(n, device="cuda", dtype=torch.float32) * sum(r + 1 for r in range(world_size))
        assert torch.equal(out, want)
        assert torch.equal(b, want), "multimem.st did not reach this replica"
    dist.barrier()
    ws.close()
    ctx.destroy()


@pytest.mark.parametrize("world", [2, 4])
def test_symmetric_workspace_peer_and_multicast(world):
    _need_gpus(world)
    from pipegoose_b200.testing.utils import spawn

    spawn(run_workspace, world_size=world)


# ------------------------------------------------------------------------------------------------ gradient kernels
def run_grad_rs_kernels(rank, world_size, port, scalar):
    import torch.distributed as dist

    os.environ["PIPEGOOSE_B200_DP_INLINE_SCALAR"] = "1" if scalar else "0"
    os.environ["PIPEGOOSE_B200_DP_INLINE_RS"] = "1"
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.ops import kernels as K
    from pipegoose_b200.ops.comm import FusedDPEngine
    from pipegoose_b200.testing.utils import init_parallel_context

    ctx = init_parallel_context(rank, world_size, port, 1, 1, world_size, backend="nccl")
    FusedDPEngine.INLINE_SCALAR_RED = scalar
    FusedDPEngine.INLINE_RS = True
    eng = FusedDPEngine(ctx, ParallelMode.DATA)
    N, Kd, M, V = 384, 256, 1024, 640            # wgrad [N, Kd] from M tokens; embedding table [V, Kd]
    head = world_size * 128                       # a small bucketed head in front of the in-kernel region
    numel = head + (N * Kd + V * Kd + 1000 + world_size * 128 - 1) // (world_size * 128) * (world_size * 128)
    _param, grad = eng.allocate(numel, torch.bfloat16, torch.float32)
    grad.zero_()
    assert eng.enable_inline(head)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    dy = (torch.randn(M, N, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    x = (torch.randn(M, Kd, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    dx = (torch.randn(M, Kd, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    ids = torch.randint(0, V, (M,), device="cuda", generator=g)
    extra_bf16 = (torch.randn(1000, device="cuda", generator=g)).to(torch.bfloat16)
    extra_f32 = torch.randn(1000, device="cuda", generator=g)
    w_view = grad[head:head + N * Kd].view(N, Kd)
    t_view = grad[head + N * Kd:head + N * Kd + V * Kd].view(V, Kd)
    e_view = grad[head + N * Kd + V * Kd:head + N * Kd + V * Kd + 1000]
    K.gemm_tn(dy, x, accum_into=w_view, accumulate=False)       # "overwrite" becomes an add at the owners
    K.gemm_tn(dy, x, accum_into=w_view, accumulate=True)        # ... and a second contribution
    K.embedding_bwd(dx, ids, V, 0, V, accum_into=t_view)
    K.accumulate_grad(extra_bf16, e_view, True)
    K.accumulate_grad(extra_f32, e_view, True, scale=0.5)
    eng.barrier()
    torch.cuda.synchronize()
    # fp32 reference of the SUM over ranks of every contribution
    ref = torch.zeros(numel, dtype=torch.float32, device="cuda")
    ref[head:head + N * Kd] = (2.0 * (dy.float().t() @ x.float())).reshape(-1)
    tab = torch.zeros(V, Kd, dtype=torch.float32, device="cuda").index_add_(0, ids, dx.float())
    ref[head + N * Kd:head + N * Kd + V * Kd] = tab.reshape(-1)
    ref[head + N * Kd + V * Kd:head + N * Kd + V * Kd + 1000] = extra_bf16.float() + 0.5 * extra_f32
    dist.all_reduce(ref)
    seg = (numel - head) // world_size
    lo, hi = head + rank * seg, head + (rank + 1) * seg
    got = grad[lo:hi]
    want = ref[lo:hi]
    err = (got - want).abs().max().item()
    assert err < 2e-2 * max(want.abs().max().item(), 1.0), err     # bf16 tensor-core products, fp32 accumulation
    # nothing may have been written locally outside the owned slice
    assert grad[head:lo].abs().max().item() == 0 if lo > head else True
    assert grad[hi:].abs().max().item() == 0 if hi < numel else True
    assert grad[:head].abs().max().item() == 0
    dist.barrier()
    eng.close_inline()
    ctx.destroy()


@pytest.mark.parametrize("world,scalar", [(2, False), (2, True), (4, False)])
def test_gradient_kernels_add_into_the_owner_slices(world, scalar):
    _need_gpus(world)
    from pipegoose_b200.testing.utils import spawn

    spawn(run_grad_rs_kernels, world_size=world, scalar=scalar)


# ------------------------------------------------------------------------------------------------ ZeRO-1 training
def run_zero_inline(rank, world_size, port, inline, nvls, state, ids, ref_losses, out_file):
    import torch.distributed as dist

    os.environ["PIPEGOOSE_B200_DP_INLINE_RS"] = "1" if inline else "0"
    os.environ["PIPEGOOSE_B200_NVLS"] = "1" if nvls else "0"
    os.environ["PIPEGOOSE_B200_NVLS_REDUCE"] = os.environ["PIPEGOOSE_B200_NVLS_ALLGATHER"] = "1" if nvls else "0"
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import DataParallel
    from pipegoose_b200.ops.comm import FusedDPEngine
    from pipegoose_b200.optim import DistributedOptimizer, FusedAdam
    from pipegoose_b200.testing.utils import init_parallel_context

    FusedDPEngine.INLINE_RS = inline
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
    for step in range(len(ref_losses)):
        if step == 1:
            # gradient accumulation over two micro-batches inside one step (no_sync on the first)
            half = local.chunk(2)
            with model.no_sync():
                l0 = model(half[0], labels=half[0]).loss
                optim.zero_grad()
                (l0 / 2).backward()
            l1 = model(half[1], labels=half[1]).loss
            (l1 / 2).backward()
            loss = (l0.detach() + l1.detach()) / 2
        else:
            loss = model(local, labels=local).loss
            optim.zero_grad()
            loss.backward()
        optim.step()
        losses.append(float(loss))
    reducer = model._pg_grad_reducer
    assert reducer._fused is not None and reducer.inline == inline
    print(f"[rank {rank}] inline {inline} workspace {reducer._fused.ws.backend}", flush=True)
    t = torch.tensor(losses, device="cuda")
    dist.all_reduce(t)
    mean_losses = (t / world_size).tolist()
    for a, b in zip(mean_losses, ref_losses):
        assert abs(a - b) < 5e-2, (mean_losses, ref_losses)
    flat = model._flat_state.flat_param.float()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "replicas diverged"
    if rank == 0:
        torch.save({"losses": mean_losses, "params": {k: v.float().cpu() for k, v in model.state_dict().items()}}, out_file)
    ctx.destroy()


@pytest.mark.parametrize("world", [2, 4])
def test_zero1_with_the_in_kernel_reduce_scatter_matches_the_bucketed_reducer(world, tmp_path):
    _need_gpus(world)
    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.optim import FusedAdam
    from pipegoose_b200.testing.utils import spawn

    torch.manual_seed(0)
    cfg = BloomConfig(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4)
    ref = BloomForCausalLM(cfg)
    state = copy.deepcopy(ref.state_dict())
    ids = torch.randint(0, 4096, (4 * world, 256))
    model = copy.deepcopy(ref).to(torch.bfloat16).cuda()
    opt = FusedAdam(model.parameters(), lr=1e-3)
    ref_losses = []
    for _ in range(3):
        loss = model(ids.cuda(), labels=ids.cuda()).loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(loss.item())
    results = {}
    for name, inline, nvls in (("inline+nvls", True, True), ("bucketed+nvls", False, True), ("bucketed", False, False)):
        out = str(tmp_path / f"{name}.pt")
        spawn(run_zero_inline, world_size=world, inline=inline, nvls=nvls, state=state, ids=ids, ref_losses=ref_losses,
              out_file=out)
        results[name] = torch.load(out)
    base = results["bucketed"]
    for name in ("inline+nvls", "bucketed+nvls"):
        for k, v in results[name]["params"].items():
            # same math, different summation order of fp32 gradients -> at most a bf16 ulp or two on a few weights
            assert torch.allclose(v, base["params"][k], atol=2e-3, rtol=2e-2), (name, k)
