"""B200-native Bloom (``BloomForCausalLM``) with the same module tree and parameter names as
🤗 transformers' Bloom, so HF state dicts load unchanged and the name-based parallel mappings
(reference nn/tensor_parallel/parallel_mapping.py:4-52) apply to both.

The forward is built from the four fused sub-layer functions in ``pipegoose_b200.ops.functional``;
activations are 2-D ``[tokens, hidden]`` bf16 tensors (token-sharded under tensor parallelism),
attention is the flash ALiBi kernel, the lm_head is fused with a vocab-parallel cross entropy
(no ``[tokens, vocab]`` fp32 tensor, no all-gather of logits).

Dropout: Bloom's ``hidden_dropout`` / ``attention_dropout`` default to 0.0, which is what the fully fused sub-layers
implement.  ``hidden_dropout > 0`` (🤗 Bloom's ``dropout_add`` after the attention projection and after the MLP) is
supported in training through a composed path: the same kernels with the residual add taken out of the GEMM epilogue
and ``dropout(.) + residual`` applied between them.  ``attention_dropout > 0`` (on the attention probabilities) takes
the same composed path with the flash kernel replaced by ``ops.attention.alibi_attention_with_dropout`` (probabilities
materialised a head group at a time).  Both only act in ``train()`` mode; ``eval()`` always runs the fused sub-layers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from pipegoose_b200.ops import functional as PF
from pipegoose_b200.ops import kernels as K


@dataclass
class BloomConfig:
    """Subset of transformers.BloomConfig that defines the architecture."""

    vocab_size: int = 250880
    hidden_size: int = 64
    n_layer: int = 2
    n_head: int = 8
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    hidden_dropout: float = 0.0
    attention_dropout: float = 0.0
    apply_residual_connection_post_layernorm: bool = False
    tie_word_embeddings: bool = True
    # "block": keep only each block's input and recompute its activations during backward (~1/3 more forward FLOPs
    # for ~n_layer x fewer saved activations); "none": save everything (default, fastest)
    recompute: str = "none"
    # architecture switches used by the other model families built from the same blocks (models/gpt2.py):
    # "alibi" (Bloom) or "learned" absolute position embeddings; LayerNorm right after the token embedding or not
    position_embedding: str = "alibi"
    n_positions: int = 0
    embedding_layernorm: bool = True

    @classmethod
    def from_hf(cls, hf_config) -> "BloomConfig":
        return cls(
            vocab_size=hf_config.vocab_size,
            hidden_size=hf_config.hidden_size,
            n_layer=hf_config.n_layer,
            n_head=hf_config.n_head,
            layer_norm_epsilon=hf_config.layer_norm_epsilon,
            initializer_range=hf_config.initializer_range,
            hidden_dropout=hf_config.hidden_dropout,
            attention_dropout=hf_config.attention_dropout,
            apply_residual_connection_post_layernorm=hf_config.apply_residual_connection_post_layernorm,
        )

    def to_hf(self):
        """The 🤗 ``transformers.BloomConfig`` of this architecture."""
        from transformers import BloomConfig as HFConfig

        return HFConfig(vocab_size=self.vocab_size, hidden_size=self.hidden_size, n_layer=self.n_layer, n_head=self.n_head,
                        layer_norm_epsilon=self.layer_norm_epsilon, initializer_range=self.initializer_range,
                        hidden_dropout=self.hidden_dropout, attention_dropout=self.attention_dropout,
                        apply_residual_connection_post_layernorm=self.apply_residual_connection_post_layernorm,
                        tie_word_embeddings=True)

    @classmethod
    def bloom_tiny(cls):
        """A toy size for CPU smoke runs of the examples."""
        return cls(vocab_size=1024, hidden_size=128, n_layer=4, n_head=8)

    @classmethod
    def bloom_560m(cls):
        return cls(hidden_size=1024, n_layer=24, n_head=16)

    @classmethod
    def bloom_1b7(cls):
        return cls(hidden_size=2048, n_layer=24, n_head=16)

    @classmethod
    def bloom_3b(cls):
        return cls(hidden_size=2560, n_layer=30, n_head=32)

    @classmethod
    def bloom_7b1(cls):
        return cls(hidden_size=4096, n_layer=30, n_head=32)


@dataclass
class CausalLMOutput:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.loss, self.logits)[i]


class BloomAttention(nn.Module):
    def __init__(self, config: BloomConfig):
        super().__init__()
        h = config.hidden_size
        self.hidden_size = h
        self.num_heads = config.n_head
        self.head_dim = h // config.n_head
        self.query_key_value = nn.Linear(h, 3 * h, bias=True)
        self.dense = nn.Linear(h, h, bias=True)
        self.use_alibi = getattr(config, "position_embedding", "alibi") == "alibi"
        self._slopes_cache = {}


class BloomMLP(nn.Module):
    def __init__(self, config: BloomConfig):
        super().__init__()
        h = config.hidden_size
        self.dense_h_to_4h = nn.Linear(h, 4 * h)
        self.dense_4h_to_h = nn.Linear(4 * h, h)

    def forward(self, hidden_states: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """HF-compatible signature ``mlp(layernorm_output, residual)`` (used by the MoE wrapper)."""
        shape = hidden_states.shape
        x = hidden_states.reshape(-1, shape[-1])
        y = PF.mlp_residual(x, self.dense_h_to_4h.weight, self.dense_h_to_4h.bias,
                            self.dense_4h_to_h.weight, self.dense_4h_to_h.bias, residual.reshape(-1, shape[-1]))
        return y.view(shape)


class BloomBlock(nn.Module):
    def __init__(self, config: BloomConfig):
        super().__init__()
        h = config.hidden_size
        self.eps = config.layer_norm_epsilon
        self.input_layernorm = nn.LayerNorm(h, eps=self.eps)
        self.self_attention = BloomAttention(config)
        self.post_attention_layernorm = nn.LayerNorm(h, eps=self.eps)
        self.mlp = BloomMLP(config)
        self.hidden_dropout = float(getattr(config, "hidden_dropout", 0.0))
        self.tp = None  # set by TensorParallel (sequence-parallel communicator)

    def _forward_with_dropout(self, x: torch.Tensor, batch: int, seq: int) -> torch.Tensor:
        """``hidden_dropout > 0`` or ``attention_dropout > 0`` in training: ``x + dropout(dense(attention))`` and
        ``x + dropout(mlp)`` — the kernels of the fused path with the residual add applied after the dropout instead of
        in the GEMM epilogue; attention probabilities are dropped in ``alibi_attention_with_dropout``."""
        import torch.nn.functional as F

        from pipegoose_b200.ops.attention import alibi_attention

        attn, tp, p = self.self_attention, self.tp, self.hidden_dropout
        n_head_local = attn.query_key_value.weight.shape[0] // (3 * attn.head_dim)
        zero = torch.zeros_like(x)
        qkv = PF.layernorm_linear(x, self.input_layernorm.weight, self.input_layernorm.bias,
                                  attn.query_key_value.weight, attn.query_key_value.bias, self.eps, tp)
        att = alibi_attention(qkv, attn.alibi_slopes_local(n_head_local), batch, seq, n_head_local, attn.head_dim,
                              dropout_p=getattr(self, "attention_dropout", 0.0))
        x = x + F.dropout(PF.linear_residual(att, attn.dense.weight, attn.dense.bias, zero, tp), p, True)
        mlp = self.mlp
        if isinstance(mlp, BloomMLP):
            h1 = K.gelu_tanh(PF.layernorm_linear(x, self.post_attention_layernorm.weight, self.post_attention_layernorm.bias,
                                                 mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias, self.eps, tp))
            return x + F.dropout(PF.linear_residual(h1, mlp.dense_4h_to_h.weight, mlp.dense_4h_to_h.bias, zero, tp), p, True)
        ln = fused_layer_norm(x, self.post_attention_layernorm.weight, self.post_attention_layernorm.bias, self.eps)
        return x + F.dropout(mlp(ln, zero), p, True)

    def forward(self, x: torch.Tensor, batch: int, seq: int) -> torch.Tensor:
        """``x``: ``[tokens_local, hidden]`` (token-sharded when ``self.tp`` is set)."""
        if self.training and (getattr(self, "hidden_dropout", 0.0) > 0.0 or getattr(self, "attention_dropout", 0.0) > 0.0):
            return self._forward_with_dropout(x, batch, seq)
        attn = self.self_attention
        tp = self.tp
        n_head_local = attn.query_key_value.weight.shape[0] // (3 * attn.head_dim)
        x = PF.attention_sublayer(x, self.input_layernorm.weight, self.input_layernorm.bias,
                                  attn.query_key_value.weight, attn.query_key_value.bias,
                                  attn.dense.weight, attn.dense.bias, attn.alibi_slopes_local(n_head_local),
                                  self.eps, batch, seq, n_head_local, attn.head_dim, tp)
        mlp = self.mlp
        if isinstance(mlp, BloomMLP):
            x = PF.layernorm_mlp(x, self.post_attention_layernorm.weight, self.post_attention_layernorm.bias,
                                 mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias,
                                 mlp.dense_4h_to_h.weight, mlp.dense_4h_to_h.bias, self.eps, tp)
        else:
            # replaced MLP (e.g. ExpertLayer): HF contract mlp(layernorm_output, residual)
            ln = fused_layer_norm(x, self.post_attention_layernorm.weight, self.post_attention_layernorm.bias, self.eps)
            x = mlp(ln, x)
        return x


def _alibi_slopes_local(self: BloomAttention, n_head_local: int) -> torch.Tensor:
    """fp32 slopes of the heads this tensor-parallel rank owns (heads are sharded contiguously);
    kept out of the module's buffers so that ``model.to(bfloat16)`` cannot round them."""
    device = self.query_key_value.weight.device
    key = (n_head_local, str(device))
    slopes = self._slopes_cache.get(key)
    if slopes is None:
        if getattr(self, "use_alibi", True):
            full = K.alibi_slopes(self.num_heads, device=device)
        else:  # no position bias inside attention (learned absolute positions): the flash kernel runs with zero slopes
            full = torch.zeros(self.num_heads, dtype=torch.float32, device=device)
        rank = getattr(self, "tp_rank", 0) if n_head_local != self.num_heads else 0
        slopes = full[rank * n_head_local:(rank + 1) * n_head_local].contiguous()
        self._slopes_cache[key] = slopes
    return slopes


BloomAttention.alibi_slopes_local = _alibi_slopes_local


class _LayerNormFn(torch.autograd.Function):
    """Stand-alone fused LayerNorm (used where the LN is not followed by one of our GEMMs)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        y, mean, rstd = K.layernorm_fwd(x2, gamma, beta, eps)
        ctx.save_for_backward(x2, gamma, beta, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma, beta, mean, rstd = ctx.saved_tensors
        dx, dg, db = PF._ln_bwd(dy.reshape(x2.shape).contiguous(), x2, gamma, beta, mean, rstd)
        return dx.view(ctx.shape), dg, db, None


def fused_layer_norm(x, gamma, beta, eps=1e-5):
    return _LayerNormFn.apply(x, gamma, beta, eps)


def left_align(attention_mask: torch.Tensor):
    """Indices that rotate every row of a ``[batch, seq]`` tensor so that its first real token comes first.

    Returns ``(idx, keep, inverse)``: ``x.gather(1, idx)`` is the left-aligned ``x`` (leading pads wrap around to the
    end), ``keep`` is the rotated mask (``True`` on real tokens) and ``y.gather(1, inverse)`` undoes the rotation.
    Masks with holes (real tokens on both sides of a pad) are refused with a device-side assert — no host sync."""
    m = attention_mask.ne(0)
    S = m.shape[1]
    lead = (m.cumsum(1) == 0).sum(1, keepdim=True)                       # pads in front of the first real token
    pos = torch.arange(S, device=m.device)[None, :]
    idx = (pos + lead) % S
    keep = m.gather(1, idx)
    torch._assert_async((keep[:, :-1] | ~keep[:, 1:]).all(),
                        "pipegoose_b200 models take left- or right-padded batches, not attention masks with holes")
    return idx, keep, (pos - lead) % S


def embed_tokens(owner: nn.Module, input_ids: torch.Tensor, config, vocab_start: int, tp) -> torch.Tensor:
    """Token ids -> ``[tokens_local, hidden]`` input of the first block.  ``owner`` holds ``word_embeddings`` and either
    ``word_embeddings_layernorm`` (Bloom) or ``position_embeddings`` (GPT-2): the model's ``transformer`` or a first
    pipeline stage."""
    if tp is not None and input_ids.numel() % tp.size != 0:
        raise ValueError(f"the sequence-parallel path shards the {input_ids.numel()} tokens of a batch (batch x seq) over "
                         f"{tp.size} tensor-parallel ranks: pad the batch so that batch x seq is a multiple of {tp.size}")
    if getattr(config, "position_embedding", "alibi") == "learned":
        assert input_ids.shape[-1] <= config.n_positions, "sequence longer than the position table"
        assert not getattr(config, "embedding_layernorm", False)
        return PF.embedding_positions(input_ids, owner.word_embeddings.weight, owner.position_embeddings.weight,
                                      vocab_start, tp)
    ln = owner.word_embeddings_layernorm
    return PF.embedding_layernorm(input_ids, owner.word_embeddings.weight, ln.weight, ln.bias,
                                  config.layer_norm_epsilon, vocab_start, tp)


def run_block(block: nn.Module, x: torch.Tensor, batch: int, seq: int, config) -> torch.Tensor:
    """One transformer block, with activation recomputation when ``config.recompute == "block"`` (training only)."""
    if getattr(config, "recompute", "none") == "block" and torch.is_grad_enabled() and x.requires_grad:
        from torch.utils.checkpoint import checkpoint

        return checkpoint(block, x, batch, seq, use_reentrant=False)
    return block(x, batch, seq)


class BloomModel(nn.Module):
    def __init__(self, config: BloomConfig):
        super().__init__()
        h = config.hidden_size
        self.config = config
        self.word_embeddings = nn.Embedding(config.vocab_size, h)
        if config.embedding_layernorm:
            self.word_embeddings_layernorm = nn.LayerNorm(h, eps=config.layer_norm_epsilon)
        if config.position_embedding == "learned":
            assert config.n_positions > 0, "learned position embeddings need n_positions"
            self.position_embeddings = nn.Embedding(config.n_positions, h)
        self.h = nn.ModuleList([BloomBlock(config) for _ in range(config.n_layer)])
        self.ln_f = nn.LayerNorm(h, eps=config.layer_norm_epsilon)


class BloomForCausalLM(nn.Module):
    base_model_prefix = "transformer"

    def __init__(self, config: BloomConfig):
        super().__init__()
        assert not config.apply_residual_connection_post_layernorm
        self.config = config
        self.transformer = BloomModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.lm_head.weight = self.transformer.word_embeddings.weight  # tied
        # the tied table receives two gradient contributions per backward (lm_head wgrad + embedding
        # scatter): the gradient reducer must wait for both before reducing its bucket
        self.lm_head.weight._pg_grad_contribs = 2
        self.tp = None
        self.vocab_start = 0
        self.apply(self._init_weights)

    def _init_weights(self, module):
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    # ------------------------------------------------------------------ helpers
    def get_input_embeddings(self):
        return self.transformer.word_embeddings

    def get_output_embeddings(self):
        return self.lm_head

    def num_parameters(self) -> int:
        seen, n = set(), 0
        for p in self.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                n += p.numel()
        return n

    def flops_per_token(self, seq_len: int) -> float:
        """Model FLOPs per token for fwd+bwd (6*N_matmul + causal attention)."""
        c = self.config
        h, L, V = c.hidden_size, c.n_layer, c.vocab_size
        per_layer = 12 * h * h
        attn = 2 * seq_len * h  # QK^T + PV, causal half of 4*S*h
        return 3 * 2 * (L * (per_layer + attn) + V * h)

    # ------------------------------------------------------------------ forward
    def hidden_states(self, input_ids: torch.Tensor) -> torch.Tensor:
        t = self.transformer
        B, S = input_ids.shape
        x = embed_tokens(t, input_ids, self.config, self.vocab_start, self.tp)
        for block in t.h:
            x = run_block(block, x, B, S, self.config)
        return x

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, **_unused) -> CausalLMOutput:
        t = self.transformer
        B, S = input_ids.shape
        unroll = None
        if attention_mask is not None:
            # The fused attention is purely causal.  RIGHT padding is harmless under a causal mask (real tokens never
            # attend to later pads).  LEFT padding (what 🤗's Bloom tokenizer produces) is turned into right padding by
            # rotating every row until its first real token sits at position 0: ALiBi only sees distances between real
            # tokens, so their hidden states are what 🤗 computes with the mask.  Pads never count in the loss.
            idx, keep, unroll = left_align(attention_mask)
            input_ids = input_ids.gather(1, idx)
            if labels is not None:
                labels = labels.gather(1, idx).masked_fill(~keep, -100)
        x = self.hidden_states(input_ids)
        eps = self.config.layer_norm_epsilon
        if labels is not None:
            # HF semantics: position s predicts token s+1; the last position has no target
            shifted = torch.full_like(labels, -100)
            shifted[:, :-1] = labels[:, 1:]
            loss = PF.lm_head_cross_entropy(x, t.ln_f.weight, t.ln_f.bias, self.lm_head.weight, shifted, eps,
                                            self.vocab_start, -100, self.tp, vocab_size=self.config.vocab_size)
            return CausalLMOutput(loss=loss, logits=None)
        ln = fused_layer_norm(x, t.ln_f.weight, t.ln_f.bias, eps)
        if self.tp is not None:
            ln = self.tp.gather_rows(ln)
        logits = PF.linear(ln, self.lm_head.weight)
        if self.tp is not None:
            logits = self.tp.gather_cols(logits)[:, : self.config.vocab_size]
        logits = logits.view(B, S, -1)
        if unroll is not None:   # back to the caller's layout (pad positions hold the logits of the rotated pads)
            logits = logits.gather(1, unroll[:, :, None].expand(-1, -1, logits.shape[-1]))
        return CausalLMOutput(loss=None, logits=logits)

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int = 1, use_cache: bool = True,
                 attention_mask: Optional[torch.Tensor] = None, do_sample: bool = False, temperature: float = 1.0,
                 top_k: int = 0, top_p: float = 1.0, eos_token_id: Optional[int] = None,
                 pad_token_id: Optional[int] = None, generator: Optional[torch.Generator] = None, **_unused) -> torch.Tensor:
        """Decoding with 🤗 ``generate``'s vocabulary for the common cases: greedy by default, ``do_sample`` with
        ``temperature`` / ``top_k`` / ``top_p``; rows stop at ``eos_token_id`` (later positions hold ``pad_token_id``,
        default: the eos id) and the loop ends when every row has stopped; returns ``input_ids`` followed by the new
        tokens.  Under tensor parallelism every rank returns the same tokens (sampled on the group's first rank).

        ``use_cache``: keys and values of every layer are kept — under tensor parallelism each rank caches the heads it
        owns — the prompt is processed once and each new token costs one position.  Mixture-of-experts blocks
        (``ExpertLayer``) decode incrementally too: the router sees the new positions only (an expert-capacity limit
        then counts the tokens of one decoding step, not of the whole sequence — with a limit that binds, cached and
        uncached decoding may drop different tokens).  🤗's left-padded prompts of unequal length (``attention_mask`` with
        leading zeros) are cached too: pad keys are masked per row, ALiBi only sees distances between real tokens.
        Masks that are not left-padded are refused (the new tokens go behind the last column).  Recomputed through the
        training forward for every token instead: models with a fused NVLink MoE layer (kernels built around
        token-sharded full sequences), ``use_cache=False``, and pipelined models (``PipelineParallel``: one forward-only
        schedule per token, the last stage picks the token and broadcasts it to the other stages)."""
        B, prompt_len = input_ids.shape
        ragged = attention_mask is not None and bool((attention_mask == 0).any())
        dense = all(isinstance(b.mlp, BloomMLP) or _decodes_incrementally(b.mlp) for b in self.transformer.h)
        group = self.tp.size if self.tp is not None else 1
        mask = attention_mask.to(input_ids.device).long() if ragged else None
        lead = None
        if ragged:
            # pads in front of each row; the KV-cache path takes LEFT-padded prompts (every row: zeros, then ones)
            lead = (mask.cumsum(1) == 0).sum(1)
            if not bool((mask.sum(1) + lead == prompt_len).all()):
                raise ValueError("generate() takes LEFT-padded prompts (attention_mask rows: zeros, then ones): new tokens "
                                 "are appended behind the last column, which must be every row's last real token")
        # a pipelined model (PipelineParallel): every new token is one forward-only schedule through the stages — keys and
        # values are not cached across stages — and the last stage, which holds the logits, tells the others the token
        engine = getattr(self, "_pg_pipeline_engine", None)
        cached = use_cache and dense and engine is None
        out = input_ids
        finished = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
        fill = pad_token_id if pad_token_id is not None else (eos_token_id if eos_token_id is not None else 0)
        cache = [None] * len(self.transformer.h)
        for step in range(max_new_tokens):
            if cached:
                last = self._incremental_logits(out if step == 0 else out[:, -1:], cache,
                                                0 if step == 0 else out.shape[1] - 1, lead)[:, -1, :]
            else:
                S = out.shape[1]
                pad = 0
                # token-sharded activations need a multiple of the group size (per micro-batch under a pipeline engine:
                # whole rows, so the row length itself is padded)
                while ((S + pad) if engine is not None else (B * (S + pad))) % group:
                    pad += 1
                if ragged:   # pads go in FRONT (mask 0): the batch stays left-padded, the last column the newest token
                    ids = out if pad == 0 else torch.cat([out.new_zeros(B, pad), out], dim=1)
                    m = mask if pad == 0 else torch.cat([mask.new_zeros(B, pad), mask], dim=1)
                    pos = -1
                else:        # right padding is harmless under a causal mask
                    ids = out if pad == 0 else torch.cat([out, out.new_zeros(B, pad)], dim=1)
                    m, pos = None, S - 1
                if engine is None:
                    last = (self(ids, attention_mask=m) if m is not None else self(ids)).logits[:, pos, :]
            if engine is not None:
                nxt = self._pipelined_next_token(engine, ids, m, pos, lambda logits: self._select_token(
                    logits, do_sample, temperature, top_k, top_p, generator))
            else:
                nxt = self._select_token(last.float(), do_sample, temperature, top_k, top_p, generator)
            if eos_token_id is not None:
                nxt = torch.where(finished, torch.full_like(nxt, fill), nxt)
                finished = finished | (nxt == eos_token_id)
            out = torch.cat([out, nxt[:, None]], dim=1)
            if ragged:
                mask = torch.cat([mask, mask.new_ones(B, 1)], dim=1)
            if eos_token_id is not None and bool(finished.all()):
                break
        return out

    def _pipelined_next_token(self, engine, ids: torch.Tensor, mask: Optional[torch.Tensor], pos: int, select) -> torch.Tensor:
        """One decoding step of a pipelined model: a forward-only schedule (micro-batched like training), the token chosen
        on the last stage from the logits at ``pos`` and broadcast over the PIPELINE group."""
        import torch.distributed as dist

        from pipegoose_b200.distributed.parallel_mode import ParallelMode

        outs = self(ids, attention_mask=mask) if mask is not None else self(ids)
        if not all(isinstance(b.mlp, BloomMLP) for b in self.transformer.h):
            # mixture-of-experts stages push their router losses on every forward: nothing consumes them while decoding
            from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext

            store = ExpertContext.get_instance()
            store.pop_all_aux_loss(), store.pop_all_z_loss()
        ctx = engine.parallel_context
        if engine.is_last:
            logits = torch.cat([o.logits if hasattr(o, "logits") else o for o in outs], dim=0)
            nxt = select(logits[:, pos, :].float())
        else:
            nxt = torch.empty(ids.shape[0], dtype=torch.long, device=ids.device)
        dist.broadcast(nxt, src=ctx.get_ranks_in_group(ParallelMode.PIPELINE)[-1], group=ctx.get_group(ParallelMode.PIPELINE))
        return nxt

    def _select_token(self, logits: torch.Tensor, do_sample: bool, temperature: float, top_k: int, top_p: float,
                      generator: Optional[torch.Generator]) -> torch.Tensor:
        """Next token of every row from its fp32 logits ``[B, V]``: arg-max, or a sample from the (temperature-scaled,
        top-k / nucleus-filtered) distribution.  Sampled tokens are made identical on every tensor-parallel rank."""
        if not do_sample:
            return logits.argmax(-1)
        if temperature != 1.0:
            logits = logits / max(float(temperature), 1e-6)
        if top_k and top_k > 0:
            kth = logits.topk(min(int(top_k), logits.shape[-1]), dim=-1).values[:, -1:]
            logits = logits.masked_fill(logits < kth, float("-inf"))
        if top_p < 1.0:
            sorted_logits, order = logits.sort(dim=-1, descending=True)
            probs = sorted_logits.softmax(-1)
            # drop a token when the mass BEFORE it already reaches top_p (the most likely token always stays)
            drop = (probs.cumsum(-1) - probs) >= top_p
            sorted_logits = sorted_logits.masked_fill(drop, float("-inf"))
            logits = torch.full_like(logits, float("-inf")).scatter(-1, order, sorted_logits)
        nxt = torch.multinomial(logits.softmax(-1), 1, generator=generator).squeeze(-1)
        if self.tp is not None and self.tp.size > 1:
            import torch.distributed as dist

            dist.broadcast(nxt, src=dist.get_global_rank(self.tp.group, 0), group=self.tp.group)
        return nxt

    def _incremental_logits(self, ids: torch.Tensor, cache: list, past: int,
                            lead: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Inference-only forward of ``ids`` (positions ``past .. past+T-1``) against the cached keys / values; plain
        tensor ops on the module's weights (the fused training kernels are built around full sequences).  ``lead[b]``:
        number of pad positions in front of row ``b`` (left-padded batch; None: no pads) — pad keys are masked with a
        large finite negative (a pad QUERY then attends uniformly instead of producing NaNs that ``0 * NaN`` would
        carry into real rows), learned absolute positions count from the first real token.  ALiBi needs no shift:
        ``slope * key_position`` differs from 🤗's mask-derived bias by a per-row constant, which softmax ignores."""
        import torch.nn.functional as F

        t, cfg = self.transformer, self.config
        B, T = ids.shape
        h, n_head = cfg.hidden_size, cfg.n_head
        D = h // n_head
        tp = self.tp
        if tp is not None:
            return self._incremental_logits_tp(ids, cache, past, lead)
        x = F.embedding(ids, t.word_embeddings.weight)
        if getattr(cfg, "position_embedding", "alibi") == "learned":
            x = x + _learned_positions(t.position_embeddings.weight, past, T, lead)
        else:
            x = F.layer_norm(x, (h,), t.word_embeddings_layernorm.weight, t.word_embeddings_layernorm.bias, cfg.layer_norm_epsilon)
        for li, block in enumerate(t.h):
            attn = block.self_attention
            ln = F.layer_norm(x, (h,), block.input_layernorm.weight, block.input_layernorm.bias, block.eps)
            qkv = F.linear(ln, attn.query_key_value.weight, attn.query_key_value.bias).view(B, T, n_head, 3, D)
            q, k, v = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))               # [B, H, T, D]
            if cache[li] is not None:
                k, v = torch.cat([cache[li][0], k], dim=2), torch.cat([cache[li][1], v], dim=2)
            cache[li] = (k, v)
            S = k.shape[2]
            scores = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(D)   # [B, H, T, S]
            key_pos = torch.arange(S, device=ids.device, dtype=torch.float32)
            if attn.use_alibi:
                scores = scores + K.alibi_slopes(n_head, device=ids.device).view(1, n_head, 1, 1) * key_pos
            query_pos = torch.arange(past, past + T, device=ids.device).view(T, 1)
            scores = scores.masked_fill(key_pos.view(1, S) > query_pos, float("-inf"))
            scores = _mask_leading_pads(scores, key_pos, lead)
            ctx = torch.matmul(scores.softmax(-1).to(v.dtype), v).transpose(1, 2).reshape(B, T, h)
            x = x + F.linear(ctx, attn.dense.weight, attn.dense.bias)
            ln = F.layer_norm(x, (h,), block.post_attention_layernorm.weight, block.post_attention_layernorm.bias, block.eps)
            mlp = block.mlp
            if not isinstance(mlp, BloomMLP):
                x = _moe_on_replicated_tokens(mlp, ln, x)
                continue
            x = x + F.linear(K.gelu_tanh(F.linear(ln, mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias)),
                             mlp.dense_4h_to_h.weight, mlp.dense_4h_to_h.bias)
        x = F.layer_norm(x[:, -1:], (h,), t.ln_f.weight, t.ln_f.bias, cfg.layer_norm_epsilon)
        return F.linear(x, self.lm_head.weight)

    def _incremental_logits_tp(self, ids: torch.Tensor, cache: list, past: int,
                               lead: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:meth:`_incremental_logits` on a tensor-parallel model: activations are replicated ``[B, T, h]`` (decoding is
        latency-bound: no sequence sharding), every rank runs the attention heads / MLP columns / vocabulary rows it owns
        and caches ITS heads' keys and values; row-parallel products and the embedding are summed over the group, the
        last position's logits are gathered along the vocabulary."""
        import torch.distributed as dist
        import torch.nn.functional as F

        t, cfg, tp = self.transformer, self.config, self.tp
        group, rank, world = tp.group, tp.rank, tp.size
        B, T = ids.shape
        h = cfg.hidden_size
        D = h // cfg.n_head

        def all_reduce(x):
            dist.all_reduce(x, group=group)
            return x

        table = t.word_embeddings.weight                      # [V_padded / world, h]
        local = ids - self.vocab_start
        mine = (local >= 0) & (local < table.shape[0])
        x = F.embedding(local.clamp(0, table.shape[0] - 1), table) * mine.unsqueeze(-1).to(table.dtype)
        x = all_reduce(x)
        if getattr(cfg, "position_embedding", "alibi") == "learned":
            x = x + _learned_positions(t.position_embeddings.weight, past, T, lead)
        else:
            x = F.layer_norm(x, (h,), t.word_embeddings_layernorm.weight, t.word_embeddings_layernorm.bias, cfg.layer_norm_epsilon)
        for li, block in enumerate(t.h):
            attn, mlp = block.self_attention, block.mlp
            n_local = attn.query_key_value.weight.shape[0] // (3 * D)
            ln = F.layer_norm(x, (h,), block.input_layernorm.weight, block.input_layernorm.bias, block.eps)
            qkv = F.linear(ln, attn.query_key_value.weight, attn.query_key_value.bias).view(B, T, n_local, 3, D)
            q, k, v = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))               # [B, H/world, T, D]
            if cache[li] is not None:
                k, v = torch.cat([cache[li][0], k], dim=2), torch.cat([cache[li][1], v], dim=2)
            cache[li] = (k, v)
            S = k.shape[2]
            scores = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(D)
            key_pos = torch.arange(S, device=ids.device, dtype=torch.float32)
            scores = scores + attn.alibi_slopes_local(n_local).view(1, n_local, 1, 1) * key_pos
            query_pos = torch.arange(past, past + T, device=ids.device).view(T, 1)
            scores = scores.masked_fill(key_pos.view(1, S) > query_pos, float("-inf"))
            scores = _mask_leading_pads(scores, key_pos, lead)
            ctx = torch.matmul(scores.softmax(-1).to(v.dtype), v).transpose(1, 2).reshape(B, T, n_local * D)
            x = x + all_reduce(F.linear(ctx, attn.dense.weight)) + attn.dense.bias      # row-parallel: bias once
            ln = F.layer_norm(x, (h,), block.post_attention_layernorm.weight, block.post_attention_layernorm.bias, block.eps)
            if not isinstance(mlp, BloomMLP):
                x = _moe_on_replicated_tokens(mlp, ln, x)    # experts sharded over the group: the layer all-reduces
                continue
            h1 = K.gelu_tanh(F.linear(ln, mlp.dense_h_to_4h.weight, mlp.dense_h_to_4h.bias))
            x = x + all_reduce(F.linear(h1, mlp.dense_4h_to_h.weight)) + mlp.dense_4h_to_h.bias
        x = F.layer_norm(x[:, -1:], (h,), t.ln_f.weight, t.ln_f.bias, cfg.layer_norm_epsilon)
        part = F.linear(x, self.lm_head.weight).contiguous()                            # [B, 1, V_padded / world]
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part, group=group)
        return torch.cat(parts, dim=-1)[..., : cfg.vocab_size]

    # ------------------------------------------------------------------ HF interop
    @classmethod
    def from_hf(cls, hf_model) -> "BloomForCausalLM":
        model = cls(BloomConfig.from_hf(hf_model.config))
        model.load_state_dict(hf_model.state_dict(), strict=False)
        return model

    def to_hf(self):
        """A 🤗 ``transformers.BloomForCausalLM`` with this model's weights (the way back from :meth:`from_hf` / the
        in-place conversion of ``TensorParallel``): parameter names are the same, so this is a config translation and a
        ``load_state_dict``.  The model must be whole — ``deparallelize()`` a live tensor- / pipeline-parallel model
        first, or merge its checkpoint with ``nn.checkpoint_convert`` and load that."""
        from transformers import BloomForCausalLM as HFBloom

        if self.tp is not None or getattr(self, "_pg_pipeline_engine", None) is not None:
            raise ValueError("to_hf() needs the unsharded model: deparallelize() first (or consolidate the checkpoint)")
        if getattr(self.config, "position_embedding", "alibi") != "alibi":
            raise ValueError("to_hf() exports Bloom-architecture models (GPT-2 family: GPT2LMHeadModel.to_hf_state_dict)")
        hf = HFBloom(getattr(self, "hf_config", None) or self.config.to_hf())
        hf = hf.to(next(self.parameters()).dtype)
        missing, unexpected = hf.load_state_dict(self.state_dict(), strict=False)
        assert not unexpected and all("lm_head" in k for k in missing), (missing, unexpected)      # (tied head)
        hf.tie_weights()
        return hf

    def save_hf_pretrained(self, path: str, **kwargs) -> None:
        """``to_hf().save_pretrained(path)``: a directory 🤗 ``from_pretrained`` reads (config.json + weights)."""
        self.to_hf().save_pretrained(path, **kwargs)


def _mask_leading_pads(scores: torch.Tensor, key_pos: torch.Tensor, lead: Optional[torch.Tensor]) -> torch.Tensor:
    """``scores [B, H, T, S]``: keys in front of row ``b``'s first real token (``key_pos < lead[b]``) get the most
    negative finite value."""
    if lead is None:
        return scores
    pad = key_pos.view(1, 1, 1, -1) < lead.view(-1, 1, 1, 1).to(key_pos.dtype)
    return scores.masked_fill(pad, torch.finfo(scores.dtype).min)


def _learned_positions(table: torch.Tensor, past: int, T: int, lead: Optional[torch.Tensor]) -> torch.Tensor:
    """Rows of a learned absolute position table for raw positions ``past .. past+T-1``; with left padding a row's
    positions count from its first real token (pads read position 0)."""
    if lead is None:
        return table[past:past + T]
    pos = (torch.arange(past, past + T, device=table.device).view(1, T) - lead.view(-1, 1)).clamp(min=0)
    return table[pos]


def _decodes_incrementally(mlp: nn.Module) -> bool:
    """Can this replaced MLP run on a few replicated positions (``mlp(layernorm_output, residual)``)?  ``ExpertLayer``
    can (router + experts on whatever tokens it is given); the fused NVLink MoE layer cannot."""
    from pipegoose_b200.nn.expert_parallel.layers import ExpertLayer

    return type(mlp) is ExpertLayer


def _moe_on_replicated_tokens(mlp: nn.Module, ln: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """One decoding step through an ``ExpertLayer``: activations are replicated over the tensor group while decoding, so
    the layer runs in its replicated-token layout (every rank routes all positions, applies the experts it owns, the
    partial outputs are all-reduced) whatever layout training uses; the router's auxiliary terms are discarded."""
    from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext

    store = ExpertContext.get_instance()
    kept_aux, kept_z = store.pop_all_aux_loss(), store.pop_all_z_loss()
    comm = getattr(mlp, "token_comm", None)
    mlp.token_comm = None
    try:
        out = mlp(ln, residual)
    finally:
        mlp.token_comm = comm
        store.pop_all_aux_loss(), store.pop_all_z_loss()
        for t in kept_aux:
            store.push_aux_loss(t)
        for t in kept_z:
            store.push_z_loss(t)
    return out


def is_hf_bloom(module: nn.Module) -> bool:
    """A 🤗 transformers ``BloomForCausalLM`` (the model the reference's README wraps)?"""
    cls = type(module)
    return cls.__name__ == "BloomForCausalLM" and cls.__module__.startswith("transformers.")


def hf_bloom_fast_path_blocker(hf_model) -> Optional[str]:
    """Why this 🤗 Bloom cannot run on the fused path (None: it can)."""
    c = hf_model.config
    if getattr(c, "apply_residual_connection_post_layernorm", False):
        return "apply_residual_connection_post_layernorm"
    if hf_model.lm_head.weight is not hf_model.transformer.word_embeddings.weight:
        return "untied lm_head"
    if c.hidden_size % c.n_head != 0:
        return "hidden_size not divisible by n_head"
    return None


def convert_hf_bloom_(hf_model) -> "BloomForCausalLM":
    """Turn a 🤗 ``BloomForCausalLM`` into this module's ``BloomForCausalLM`` IN PLACE: the same Python objects, the
    same ``nn.Parameter``s under the same names (optimizers, state dicts and checkpoints written from either side keep
    working) — only the container classes change, so that ``forward`` runs the fused sub-layer kernels (flash ALiBi
    attention, GEMMs with fused epilogues, fused lm_head + cross entropy) and ``TensorParallel`` can take the
    sequence-parallel path with the fused all-gather->GEMM / GEMM->reduce-scatter kernels.

    This is what ``TensorParallel(hf_bloom, ctx).parallelize()`` does for the reference's canonical input
    (reference README.md:21-69, examples/hybrid_parallelism.py:22-37)."""
    why = hf_bloom_fast_path_blocker(hf_model)
    if why is not None:
        raise ValueError(f"this Bloom cannot use the fused path: {why}")
    cfg = BloomConfig.from_hf(hf_model.config)
    cfg.tie_word_embeddings = True
    hf_config = hf_model.config
    t = hf_model.transformer
    def is_hf(mod, name):
        return type(mod).__name__ == name and type(mod).__module__.startswith("transformers.")

    for block in t.h:
        attn = block.self_attention
        attn._modules.pop("attention_dropout", None)
        attn.__class__ = BloomAttention
        attn.hidden_size, attn.num_heads, attn.head_dim = cfg.hidden_size, cfg.n_head, cfg.hidden_size // cfg.n_head
        attn.use_alibi = True
        attn._slopes_cache = {}
        # the block's MLP, or — when ExpertParallel already replaced it by an ExpertLayer — the experts inside it
        # (the ``mlp(layernorm_output, residual)`` contract is the same for both classes)
        for mlp in [m for m in block.modules() if is_hf(m, "BloomMLP")]:
            mlp._modules.pop("gelu_impl", None)
            mlp.__class__ = BloomMLP
        block.__class__ = BloomBlock
        block.eps = cfg.layer_norm_epsilon
        block.hidden_dropout = float(cfg.hidden_dropout)
        block.attention_dropout = float(cfg.attention_dropout)
        block.tp = None
    t.__class__ = BloomModel
    t.config = cfg
    hf_model.__class__ = BloomForCausalLM
    hf_model.config = cfg
    hf_model.hf_config = hf_config          # what the user built the model from (save / export helpers can read it)
    hf_model.tp = None
    hf_model.vocab_start = 0
    hf_model.lm_head.weight._pg_grad_contribs = 2   # tied table: lm_head wgrad + embedding backward
    return hf_model
