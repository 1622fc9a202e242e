"""Thin functional wrappers over the native kernels with PyTorch reference fallbacks.

Each wrapper has one contract and two implementations: the sm_100a kernel (CUDA bf16 tensors,
mandatory on GPU machines) and plain PyTorch math (CPU / fp32, used by the gloo unit tests and
as the numerics oracle in the GPU tests).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from pipegoose_b200.ops import native, use_native

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_OUT_F32, EPI_ACCUM, EPI_DGELU, EPI_SCATTER = 1, 2, 4, 8, 16, 32, 64


# ----------------------------------------------------------------------------------------------
# in-kernel data-parallel gradient reduce-scatter (ZeRO-1): address ranges of flat fp32 gradient buffers whose
# contributions are added straight into the OWNER rank's buffer by the producing kernel (csrc/grad_rs.cuh).  Every
# gradient writer below looks its destination up here, so a writer can never bypass the reduction by accident.
# ----------------------------------------------------------------------------------------------
_GRAD_RS_REGIONS = []   # [(first byte, end byte, descriptor dict for the kernels, engine)]


def register_grad_rs(begin: int, end: int, desc: dict, engine) -> None:
    import weakref

    unregister_grad_rs(engine)
    _GRAD_RS_REGIONS.append((begin, end, desc, weakref.ref(engine)))   # (weak: a dropped engine unregisters itself)


def unregister_grad_rs(engine) -> None:
    _GRAD_RS_REGIONS[:] = [r for r in _GRAD_RS_REGIONS if r[3]() is not None and r[3]() is not engine]


def grad_rs_for(t: Optional[torch.Tensor]):
    """The reduce-scatter descriptor if ``t`` (a view of a flat gradient buffer) lies in a registered region."""
    if not _GRAD_RS_REGIONS or t is None or not t.is_cuda:
        return None
    ptr = t.data_ptr()
    for begin, end, desc, ref in _GRAD_RS_REGIONS:
        engine = ref()
        if engine is not None and begin <= ptr < end:
            engine.inline_dirty = True
            return desc
    return None


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """Bloom's GELU (tanh approximation; transformers BloomGelu)."""
    return x * 0.5 * (1.0 + torch.tanh(0.79788456 * x * (1.0 + 0.044715 * x * x)))


def gelu_tanh_grad(x: torch.Tensor) -> torch.Tensor:
    t = torch.tanh(0.79788456 * x * (1.0 + 0.044715 * x * x))
    return 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)


# ----------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------
def gemm_nt(x, w, bias=None, residual=None, gelu=False, aux_out=None, out=None, **kw):
    """``out[M,N] = epi(x[M,K] @ w[N,K]^T)``: bias, tanh-GELU (pre-activation into ``aux_out``), residual."""
    if use_native(x, w):
        M, N = x.shape[0], w.shape[0]
        if out is None:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
        flags = EPI_GELU if gelu else 0
        native().gemm(x, w, out, False, False, bias, residual, aux_out if gelu else None, flags, **kw)
        return out
    y = x @ w.t()
    if bias is not None:
        y = y + bias
    if gelu:
        if aux_out is not None:
            aux_out.copy_(y)
        y = gelu_tanh(y)
    if residual is not None:
        y = y + residual
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_nn(dy, w, dgelu_aux=None, out=None, **kw):
    """``out[M,K] = dy[M,N] @ w[N,K]`` (dgrad); optional ``* gelu'(aux)`` epilogue."""
    if use_native(dy, w):
        M, K = dy.shape[0], w.shape[1]
        if out is None:
            out = torch.empty(M, K, dtype=torch.bfloat16, device=dy.device)
        flags = EPI_DGELU if dgelu_aux is not None else 0
        native().gemm(dy, w, out, False, True, None, None, dgelu_aux, flags, **kw)
        return out
    y = dy @ w
    if dgelu_aux is not None:
        y = y * gelu_tanh_grad(dgelu_aux.float()).to(y.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_tn(dy, x, accum_into: Optional[torch.Tensor] = None, accumulate: bool = True, **kw):
    """``dW[N,K] = dy[M,N]^T @ x[M,K]`` (wgrad).  With ``accum_into`` (fp32 main grad) the product is
    accumulated in place by the GEMM epilogue and ``None`` is returned."""
    if use_native(dy, x):
        N, K = dy.shape[1], x.shape[1]
        if accum_into is not None:
            rs = grad_rs_for(accum_into)
            if rs is not None:
                # wgrad -> reduce-scatter in one kernel: the epilogue adds every tile into its owner's gradient buffer
                kw = dict(kw)
                kw["ag"] = dict(kw.get("ag") or {}, grad_rs=rs)
                accumulate = True
            native().gemm(dy, x, accum_into.view(N, K), True, True, None, None, None,
                          EPI_ACCUM if accumulate else 0, **kw)
            return None
        out = torch.empty(N, K, dtype=torch.bfloat16, device=dy.device)
        native().gemm(dy, x, out, True, True, None, None, None, 0, **kw)
        return out
    g = dy.t() @ x
    if accum_into is not None:
        if accumulate:
            accum_into.view_as(g).add_(g.to(accum_into.dtype))
        else:
            accum_into.view_as(g).copy_(g.to(accum_into.dtype))
        return None
    return g


def colsum(dy, accum_into: Optional[torch.Tensor] = None):
    """Bias gradient: column sums of ``dy`` (accumulated into the fp32 main grad when given)."""
    if use_native(dy):
        if accum_into is not None:
            native().colsum(dy, accum_into)
            return None
        out = torch.zeros(dy.shape[1], dtype=torch.float32, device=dy.device)
        native().colsum(dy, out)
        return out.to(dy.dtype)
    g = dy.float().sum(0)
    if accum_into is not None:
        accum_into.add_(g.to(accum_into.dtype))
        return None
    return g.to(dy.dtype)


def accumulate_grad(grad, main_grad, accumulate: bool = True, scale: float = 1.0):
    """``main_grad (+)= scale * grad``: cast (bf16 -> fp32), scale and accumulate in one pass."""
    rs = grad_rs_for(main_grad)
    if rs is not None:
        # the destination is reduce-scattered in-kernel: the contribution goes to the owners of its slices
        src = grad if grad.dtype in (torch.bfloat16, torch.float32) else grad.float()
        native().grad_rs_accum(src.contiguous().view(-1), main_grad.view(-1), scale, rs)
        return
    if use_native(grad) and main_grad.dtype == torch.float32 and grad.is_contiguous():
        native().accum_bf16_to_f32(grad, main_grad, scale, accumulate)
        return
    g = grad.to(main_grad.dtype) * scale if scale != 1.0 else grad.to(main_grad.dtype)
    if accumulate:
        main_grad.add_(g.view_as(main_grad))
    else:
        main_grad.copy_(g.view_as(main_grad))


# ----------------------------------------------------------------------------------------------
# LayerNorm (+ fused embedding gather)
# ----------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, ids=None, vocab_start=0, vocab_end=0, apply_ln=True, out=None):
    """Returns ``(y, mean, rstd)``.  With ``ids`` the input rows are gathered from the table ``x``.
    ``out``: optional preallocated bf16 destination (e.g. a peer-visible all-gather staging buffer)."""
    if use_native(x):
        h = x.shape[-1]
        rows = ids.numel() if ids is not None else x.numel() // h
        y = out if out is not None else torch.empty(rows, h, dtype=torch.bfloat16, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        native().layernorm_fwd(x, ids.reshape(-1) if ids is not None else None, vocab_start, vocab_end,
                               gamma, beta, y, mean, rstd, eps, apply_ln)
        return y, mean, rstd
    if ids is not None:
        flat = ids.reshape(-1)
        mask = (flat >= vocab_start) & (flat < vocab_end)
        src = x[(flat - vocab_start).clamp(0, x.shape[0] - 1)] * mask.unsqueeze(-1).to(x.dtype)
    else:
        src = x.reshape(-1, x.shape[-1])
    if not apply_ln:
        return src, None, None
    xf = src.float()
    mean = xf.mean(-1)
    var = xf.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = ((xf - mean[:, None]) * rstd[:, None] * gamma.float() + beta.float()).to(src.dtype)
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_extra=None, dgamma_acc=None, dbeta_acc=None, dx_out=None):
    """Returns ``(dx, dgamma, dbeta)``; the parameter grads are ``None`` when accumulated in place.  ``dx_out``:
    write dx into this buffer (the next all-gather's staging slot) instead of a fresh tensor."""
    if use_native(dy):
        dx = dx_out if dx_out is not None else torch.empty_like(x)
        if dgamma_acc is None:
            dg = torch.zeros(x.shape[-1], dtype=torch.float32, device=x.device)
            db = torch.zeros_like(dg)
            native().layernorm_bwd(dy, x, gamma, mean, rstd, dx_extra, dx, dg, db)
            return dx, dg.to(gamma.dtype), db.to(gamma.dtype)
        native().layernorm_bwd(dy, x, gamma, mean, rstd, dx_extra, dx, dgamma_acc, dbeta_acc)
        return dx, None, None
    xf, dyf = x.float(), dy.float()
    xhat = (xf - mean[:, None]) * rstd[:, None]
    g = dyf * gamma.float()
    s1 = g.mean(-1, keepdim=True)
    s2 = (g * xhat).mean(-1, keepdim=True)
    dx = rstd[:, None] * (g - s1 - xhat * s2)
    if dx_extra is not None:
        dx = dx + dx_extra.float()
    dgamma = (dyf * xhat).sum(0)
    dbeta = dyf.sum(0)
    if dgamma_acc is not None:
        dgamma_acc.add_(dgamma.to(dgamma_acc.dtype))
        dbeta_acc.add_(dbeta.to(dbeta_acc.dtype))
        return dx.to(x.dtype), None, None
    return dx.to(x.dtype), dgamma.to(gamma.dtype), dbeta.to(gamma.dtype)


def embedding_bwd(dx, ids, vocab_rows, vocab_start, vocab_end, accum_into=None):
    if use_native(dx):
        tgt = accum_into if accum_into is not None else torch.zeros(vocab_rows, dx.shape[-1], dtype=torch.float32, device=dx.device)
        native().embedding_bwd(dx, ids.reshape(-1), tgt, vocab_start, vocab_end, grad_rs_for(accum_into) or {})
        return None if accum_into is not None else tgt.to(dx.dtype)
    flat = ids.reshape(-1)
    mask = (flat >= vocab_start) & (flat < vocab_end)
    g = torch.zeros(vocab_rows, dx.shape[-1], dtype=torch.float32, device=dx.device)
    g.index_add_(0, (flat - vocab_start)[mask], dx.float()[mask])
    if accum_into is not None:
        accum_into.view_as(g).add_(g.to(accum_into.dtype))
        return None
    return g.to(dx.dtype)


# ----------------------------------------------------------------------------------------------
# vocab-parallel cross entropy
# ----------------------------------------------------------------------------------------------
def ce_local_stats(logits, targets, vocab_start):
    """Per row ``(max, sum exp(x - max), target logit or 0)`` over this rank's vocab shard -> [rows, 3] fp32."""
    if use_native(logits):
        stats = torch.empty(logits.shape[0], 3, dtype=torch.float32, device=logits.device)
        native().ce_stats(logits, targets, stats, vocab_start)
        return stats
    lf = logits.float()
    m = lf.max(-1).values
    s = torch.exp(lf - m[:, None]).sum(-1)
    t = targets - vocab_start
    ok = (t >= 0) & (t < logits.shape[1])
    tl = torch.where(ok, lf.gather(1, t.clamp(0, logits.shape[1] - 1)[:, None]).squeeze(1), torch.zeros_like(m))
    return torch.stack([m, s, tl], dim=1)


def ce_partials_buffer(rows: int, vocab_local: int, device) -> torch.Tensor:
    """Buffer for the online-softmax partials the lm_head GEMM epilogue emits: two (max, sumexp) pairs per row and
    256-column tile."""
    return torch.empty(rows, 2 * ((vocab_local + 255) // 256), 2, dtype=torch.float32, device=device)


def ce_stats_from_partials(part, logits, targets, vocab_start):
    """``[rows, 3]`` (max, sum exp(x - max), target logit or 0) from the GEMM epilogue's partials — the statistics pass
    over the ``[rows, vocab]`` logits is gone, only ``rows x tiles x 16`` bytes are read."""
    stats = torch.empty(logits.shape[0], 3, dtype=torch.float32, device=logits.device)
    native().ce_combine(part, logits, targets, stats, vocab_start)
    return stats


def ce_combine_stats(all_stats: torch.Tensor) -> torch.Tensor:
    """Merge ``[T, rows, 3]`` per-shard stats into global ``[rows, 3]`` (max, sumexp, target logit)."""
    m = all_stats[..., 0].max(0).values
    s = (all_stats[..., 1] * torch.exp(all_stats[..., 0] - m)).sum(0)
    tl = all_stats[..., 2].sum(0)
    return torch.stack([m, s, tl], dim=1).contiguous()


def ce_finalize(logits, targets, gstats, vocab_start, grad_scale, ignore_index=-100, write_grad=True):
    """Returns per-row loss; overwrites ``logits`` with ``dlogits * grad_scale`` when ``write_grad``.
    ``grad_scale`` is a 1-element fp32 device tensor (no host sync) or ``None`` (= 1)."""
    if use_native(logits):
        loss_rows = torch.empty(logits.shape[0], dtype=torch.float32, device=logits.device)
        native().ce_finalize(logits, targets, gstats, loss_rows, vocab_start, grad_scale, ignore_index, write_grad)
        return loss_rows
    grad_scale = 1.0 if grad_scale is None else grad_scale.float()
    gm, gs, tl = gstats[:, 0], gstats[:, 1], gstats[:, 2]
    ignored = targets == ignore_index
    loss = torch.where(ignored, torch.zeros_like(gm), torch.log(gs) + gm - tl)
    if write_grad:
        p = torch.exp(logits.float() - gm[:, None]) / gs[:, None]
        t = targets - vocab_start
        ok = (t >= 0) & (t < logits.shape[1]) & ~ignored
        rows = torch.nonzero(ok).squeeze(1)
        p[rows, t[rows]] -= 1.0
        p = p * grad_scale
        p[ignored] = 0
        logits.copy_(p.to(logits.dtype))
    return loss


# ----------------------------------------------------------------------------------------------
# ALiBi causal attention (reference math; the tcgen05 kernel lives in ops/attention.py)
# ----------------------------------------------------------------------------------------------
def alibi_slopes(n_head: int, device=None) -> torch.Tensor:
    """Per-head ALiBi slopes exactly as transformers' ``build_alibi_tensor`` computes them."""
    closest = 2 ** math.floor(math.log2(n_head))
    base = 2.0 ** (-(2.0 ** -(math.log2(closest) - 3)))
    slopes = torch.pow(torch.tensor(base, dtype=torch.float32), torch.arange(1, 1 + closest, dtype=torch.float32))
    if closest != n_head:
        extra_base = 2.0 ** (-(2.0 ** -(math.log2(2 * closest) - 3)))
        n_rem = min(closest, n_head - closest)
        extra = torch.pow(torch.tensor(extra_base, dtype=torch.float32), torch.arange(1, 1 + 2 * n_rem, 2, dtype=torch.float32))
        slopes = torch.cat([slopes, extra])
    return slopes.to(device)
