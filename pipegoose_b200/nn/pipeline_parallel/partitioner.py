"""Structural pipeline partitioning (parity: reference nn/pipeline_parallel/partitioner.py:29-244).

The reference symbolically traces the model with ``transformers.utils.fx`` (removed in
transformers 5) and cuts the FX graph at transformer-block boundaries once a shard holds its share
of the (non-embedding) parameters.  This partitioner reaches the same cuts without tracing: it
reads the model's block list (``<base_model_prefix>.h`` / ``.layers`` or the children of an
``nn.Sequential``), balances blocks by parameter count, and returns one ``nn.Module`` per stage
whose forward consumes the previous stage's hidden states.
"""
from __future__ import annotations

from enum import Enum, auto
from typing import List, Optional

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext


class PartitionPolicy(Enum):
    UNIFORM = auto()


def _balanced_cuts(costs: List[int], n_parts: int) -> List[int]:
    """Contiguous split of ``costs`` into ``n_parts`` groups minimising the largest group (greedy on the
    running prefix, same rule as the reference: move to the next shard once its share is reached)."""
    total = sum(costs)
    bounds, acc, part = [0], 0, 1
    for i, c in enumerate(costs):
        remaining_items = len(costs) - i
        remaining_parts = n_parts - part + 1
        if part < n_parts and (acc >= total * part / n_parts or remaining_items < remaining_parts) and i > bounds[-1]:
            bounds.append(i)
            part += 1
        acc += c
    while len(bounds) < n_parts:
        bounds.append(min(bounds[-1] + 1, len(costs)))
    bounds.append(len(costs))
    return bounds


class SequentialStage(nn.Module):
    def __init__(self, layers: List[nn.Module]):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


_HF_BLOOM_BOOL_MASK = None


def _hf_bloom_uses_boolean_mask() -> bool:
    """Old 🤗 Bloom blocks take a boolean mask (``masked_fill``), newer ones an additive float mask."""
    global _HF_BLOOM_BOOL_MASK
    if _HF_BLOOM_BOOL_MASK is None:
        import inspect

        from transformers.models.bloom.modeling_bloom import BloomAttention

        _HF_BLOOM_BOOL_MASK = "masked_fill" in inspect.getsource(BloomAttention.forward)
    return _HF_BLOOM_BOOL_MASK


class _ReferenceCallStyle:
    """The reference's partitions are chained like ``out = stage0(**inputs); out = stage1(*out); ...`` (its
    tests/nn/pipeline_parallel/test_partitioner.py): a first stage called with ``input_ids=...`` — or a later one called
    with the positional tuple a previous stage returned — answers in that style: every stage but the last returns
    ``(hidden, attention_mask, None, (batch, seq))``, i.e. the positional arguments of the next stage.  The engines of
    this library call ``stage(x, attention_mask=..., batch_seq=...)`` and get plain tensors."""

    def __call__(self, *args, input_ids=None, **kwargs):
        chained = input_ids is not None or len(args) >= 2
        if input_ids is not None:
            args = (input_ids,) + args
        out = super().__call__(*args, **kwargs)
        if not chained or self.is_last:
            return out
        mask = kwargs.get("attention_mask", args[1] if len(args) > 1 else None)
        batch_seq = kwargs.get("batch_seq", args[3] if len(args) > 3 else None)
        if batch_seq is None and self.is_first:
            batch_seq = tuple(args[0].shape[:2])
        return (out, mask, None, batch_seq)


class BloomStage(_ReferenceCallStyle, nn.Module):
    """A contiguous slice of a Bloom-style causal LM: [embedding +] blocks [+ final norm + lm head].

    Works for ``pipegoose_b200.models.BloomForCausalLM`` (fused blocks on 2-D token tensors) and for
    🤗 ``BloomForCausalLM`` (blocks called with alibi / causal mask)."""

    def __init__(self, model: nn.Module, start: int, end: int, is_first: bool, is_last: bool):
        super().__init__()
        t = model.transformer
        self.is_first, self.is_last = is_first, is_last
        self.config = model.config
        self.fast = hasattr(model, "hidden_states")  # our fused model
        if is_first:
            self.word_embeddings = t.word_embeddings
            for name in ("word_embeddings_layernorm", "position_embeddings"):  # Bloom / GPT-2 style front end
                if hasattr(t, name):
                    setattr(self, name, getattr(t, name))
        self.h = nn.ModuleList([t.h[i] for i in range(start, end)])
        if is_last:
            self.ln_f = t.ln_f
            self.lm_head = model.lm_head
        self._model_ref = [model]  # not registered: avoids duplicating parameters

    def count_targets(self, labels: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """How many positions of this (micro-)batch the loss is averaged over (for the engine's micro-batch weights).
        The fused stages never score pads, and a row starts with its first REAL token, which has no predecessor."""
        if not self.fast or attention_mask is None:
            return (labels[..., 1:] != -100).sum()
        from pipegoose_b200.models.bloom import left_align

        idx, keep, _ = left_align(attention_mask)
        aligned = labels.gather(1, idx).masked_fill(~keep, -100)
        return (aligned[:, 1:] != -100).sum()

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, batch_seq=None):
        model = self._model_ref[0]
        if self.fast:
            from pipegoose_b200.models.bloom import fused_layer_norm
            from pipegoose_b200.ops import functional as PF
            from pipegoose_b200.ops import kernels as K

            eps = self.config.layer_norm_epsilon
            idx = keep = unroll = None
            if attention_mask is not None and (self.is_first or self.is_last):
                # left padding -> right padding, pads out of the loss (models/bloom.py::left_align); every stage derives
                # the same rotation from the mask, the blocks in between only see token rows
                from pipegoose_b200.models.bloom import left_align

                idx, keep, unroll = left_align(attention_mask)
            if self.is_first:
                B, S = x.shape
                from pipegoose_b200.models.bloom import embed_tokens

                h = embed_tokens(self, x if idx is None else x.gather(1, idx), self.config, model.vocab_start, model.tp)
            else:
                B, S = batch_seq
                h = x
            from pipegoose_b200.models.bloom import run_block

            for block in self.h:
                h = run_block(block, h, B, S, self.config)
            if not self.is_last:
                return h
            if labels is not None:
                if idx is not None:
                    labels = labels.gather(1, idx).masked_fill(~keep, -100)
                shifted = torch.full_like(labels, -100)
                shifted[:, :-1] = labels[:, 1:]
                return PF.lm_head_cross_entropy(h, self.ln_f.weight, self.ln_f.bias, self.lm_head.weight, shifted,
                                                eps, model.vocab_start, -100, model.tp,
                                                vocab_size=self.config.vocab_size)
            ln = fused_layer_norm(h, self.ln_f.weight, self.ln_f.bias, eps)
            # as the unpartitioned model (models/bloom.py): under tensor parallelism the rows are token-sharded and the
            # table vocabulary-sharded — gather the tokens, multiply, gather the vocabulary; ``PF.linear`` (not the bare
            # kernel) so that ``out.sum().backward()`` of the reference's forward-only usage reaches the lm_head
            tp = model.tp
            if tp is not None:
                ln = tp.gather_rows(ln)
            logits = PF.linear(ln, self.lm_head.weight)
            if tp is not None:
                logits = tp.gather_cols(logits)[:, : self.config.vocab_size]
            logits = logits.view(B, S, -1)
            if unroll is not None:
                logits = logits.gather(1, unroll[:, :, None].expand(-1, -1, logits.shape[-1]))
            return logits
        # ---- 🤗 Bloom blocks
        from transformers.models.bloom.modeling_bloom import build_alibi_tensor

        if self.is_first:
            ids = x
            B, S = ids.shape
            h = self.word_embeddings_layernorm(self.word_embeddings(ids))
        else:
            h = x
            B, S = h.shape[:2]
        if attention_mask is None:
            attention_mask = torch.ones(B, S, dtype=torch.long, device=h.device)
        alibi = build_alibi_tensor(attention_mask, self.config.n_head, dtype=h.dtype)
        causal = torch.ones(S, S, dtype=torch.bool, device=h.device).triu(1)[None, None].expand(B, 1, S, S)
        if not _hf_bloom_uses_boolean_mask():
            # newer transformers ADD the mask to the scores: 0 where visible, a large negative where masked
            causal = torch.zeros(B, 1, S, S, dtype=h.dtype, device=h.device).masked_fill(causal, torch.finfo(h.dtype).min)
        for block in self.h:
            out = block(h, alibi=alibi, attention_mask=causal)
            h = out[0] if isinstance(out, tuple) else out
        if not self.is_last:
            return h
        logits = self.lm_head(self.ln_f(h))
        if labels is not None:
            shift_logits = logits[..., :-1, :].contiguous().float()
            shift_labels = labels[..., 1:].contiguous()
            return torch.nn.functional.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1))
        return logits


class GPT2Stage(_ReferenceCallStyle, nn.Module):
    """A contiguous slice of a 🤗 GPT-2 style causal LM (``transformer.{wte,wpe,drop,h,ln_f}`` + ``lm_head``) —
    the second model family the reference's partitioner is tested with (tests/nn/pipeline_parallel/test_partitioner.py)."""

    def __init__(self, model: nn.Module, start: int, end: int, is_first: bool, is_last: bool):
        super().__init__()
        t = model.transformer
        self.is_first, self.is_last = is_first, is_last
        if is_first:
            self.wte, self.wpe, self.drop = t.wte, t.wpe, t.drop
        self.h = nn.ModuleList([t.h[i] for i in range(start, end)])
        if is_last:
            self.ln_f = t.ln_f
            self.lm_head = model.lm_head

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, batch_seq=None):
        if self.is_first:
            pos = torch.arange(x.shape[1], device=x.device)[None]
            h = self.drop(self.wte(x) + self.wpe(pos))
        else:
            h = x
        for block in self.h:
            out = block(h)
            h = out[0] if isinstance(out, tuple) else out
        if not self.is_last:
            return h
        logits = self.lm_head(self.ln_f(h))
        if labels is not None:
            shift_logits = logits[..., :-1, :].contiguous().float()
            shift_labels = labels[..., 1:].contiguous()
            return torch.nn.functional.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1))
        return logits


class RotaryDecoderStage(_ReferenceCallStyle, nn.Module):
    """A contiguous slice of a 🤗 LLaMA-style causal LM (``model.{embed_tokens,layers,norm,rotary_emb}`` + ``lm_head``:
    LLaMA, Mistral, Qwen2, ... — pre-norm blocks with rotary position embeddings).  Every stage recomputes the rotary
    tables from its own (parameter-free) ``rotary_emb`` and builds the additive causal mask, so only hidden states
    travel between stages."""

    def __init__(self, model: nn.Module, start: int, end: int, is_first: bool, is_last: bool):
        super().__init__()
        base = model.model
        self.is_first, self.is_last = is_first, is_last
        if is_first:
            self.embed_tokens = base.embed_tokens
        self.layers = nn.ModuleList([base.layers[i] for i in range(start, end)])
        self.rotary_emb = base.rotary_emb
        if is_last:
            self.norm = base.norm
            self.lm_head = model.lm_head

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, batch_seq=None):
        h = self.embed_tokens(x) if self.is_first else x
        B, S = h.shape[:2]
        pos = torch.arange(S, device=h.device)[None]
        rope = self.rotary_emb(h, pos)
        hidden = torch.ones(S, S, dtype=torch.bool, device=h.device).triu(1)[None, None]
        if attention_mask is not None:  # padding: keys of padded positions are invisible
            hidden = hidden | (attention_mask[:, None, None, :] == 0)
        mask = torch.zeros(B, 1, S, S, dtype=h.dtype, device=h.device).masked_fill(hidden, torch.finfo(h.dtype).min)
        for layer in self.layers:
            out = layer(h, attention_mask=mask, position_ids=pos, position_embeddings=rope)
            h = out[0] if isinstance(out, tuple) else out
        if not self.is_last:
            return h
        logits = self.lm_head(self.norm(h))
        if labels is not None:
            shift_logits = logits[..., :-1, :].contiguous().float()
            shift_labels = labels[..., 1:].contiguous()
            return torch.nn.functional.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1))
        return logits


# Names of the tensors a pipeline's first stage takes (parity: reference partitioner.py:14); ``split`` accepts the
# argument for call compatibility, the stage modules here take these two by keyword.
INPUT_NAMES = ["input_ids", "attention_mask"]


class BasePartitioner:
    """``split(input_names) -> [stage modules]`` (parity: reference partitioner.py:20-26)."""

    def split(self, input_names: Optional[List[str]] = None) -> List[nn.Module]:
        raise NotImplementedError


class UniformPartitioner(BasePartitioner):
    # Model families beyond the built-in ones (nn.Sequential, Bloom-, GPT-2- and LLaMA-style causal LMs):
    # [(matches(model) -> bool, blocks(model) -> list of blocks, stage_cls(model, start, end, is_first, is_last))].
    # The reference reaches arbitrary 🤗 models by tracing them with transformers.utils.fx (partitioner.py:146-219);
    # transformers 5 removed that tracer, so a family this partitioner does not know is described here instead, the way
    # TensorParallelMapping.register describes a new family to the tensor-parallel wrapper.
    _FAMILIES: List = []

    @classmethod
    def register_family(cls, matches, blocks, stage_cls) -> None:
        """Teach the partitioner a new model family.  ``matches(model)`` recognises it, ``blocks(model)`` returns its
        transformer blocks in order (they are balanced by parameter count), and ``stage_cls(model, start, end,
        is_first, is_last)`` builds the ``nn.Module`` that runs blocks ``[start, end)`` — plus the embedding front end
        when ``is_first`` and the final norm / head / loss when ``is_last`` — with the signature
        ``forward(x, attention_mask=None, labels=None, batch_seq=None)`` of the built-in stages."""
        cls._FAMILIES.insert(0, (matches, blocks, stage_cls))

    def __init__(self, model: nn.Module, parallel_context: ParallelContext):
        self.module = model
        self.parallel_context = parallel_context

    def _n_partitions(self) -> int:
        return self.parallel_context.pipeline_parallel_size

    @staticmethod
    def _block_list(model: nn.Module) -> Optional[nn.ModuleList]:
        base = getattr(model, getattr(model, "base_model_prefix", "transformer"), None)
        if base is None:
            return None
        for name in ("h", "layers", "layer", "blocks"):
            blocks = getattr(base, name, None)
            if isinstance(blocks, nn.ModuleList):
                return blocks
        return None

    def split(self, input_names: Optional[List[str]] = None) -> List[nn.Module]:
        n = self._n_partitions()
        model = self.module
        if isinstance(model, nn.Sequential):
            layers = list(model.children())
            costs = [sum(p.numel() for p in l.parameters()) or 1 for l in layers]
            b = _balanced_cuts(costs, n)
            return [SequentialStage(layers[b[i]:b[i + 1]]) for i in range(n)]
        for matches, family_blocks, family_stage in self._FAMILIES:
            if matches(model):
                blocks = list(family_blocks(model))
                assert len(blocks) >= n, "more pipeline stages than transformer blocks"
                b = _balanced_cuts([sum(p.numel() for p in blk.parameters()) or 1 for blk in blocks], n)
                return [family_stage(model, b[i], b[i + 1], i == 0, i == n - 1) for i in range(n)]
        blocks = self._block_list(model)
        rotary = blocks is not None and hasattr(getattr(model, "model", None), "rotary_emb") and hasattr(model, "lm_head")
        if blocks is None or not (hasattr(model, "transformer") or rotary):
            # not a family this partitioner knows by structure: cut its torch.fx graph where one activation is live
            from pipegoose_b200.nn.pipeline_parallel.fx_partitioner import GraphPartitioner, NoLegalCut

            why = ""
            try:
                try:
                    return GraphPartitioner(model, self.parallel_context, n_partitions=n).split(input_names)
                except NoLegalCut:   # long skip connections: let up to four activations cross a cut (packed into one buffer)
                    return GraphPartitioner(model, self.parallel_context, n_partitions=n, max_boundary_tensors=4).split(input_names)
            except NoLegalCut as e:
                why = f"  Its torch.fx graph cannot be cut either: {e}"
            except Exception as e:   # torch.fx could not trace it (data-dependent control flow, ...)
                why = f"  torch.fx could not trace it either ({type(e).__name__}: {str(e).splitlines()[0][:200] if str(e) else ''})."
            raise NotImplementedError(
                f"UniformPartitioner does not know how to cut a {type(model).__name__} into pipeline stages.  Built in: "
                "nn.Sequential, Bloom-style (transformer.h + word_embeddings), GPT-2-style (transformer.wte/wpe/h) and "
                "LLaMA-style (model.embed_tokens/layers/norm/rotary_emb + lm_head) causal LMs.  The reference traces any "
                "🤗 model with transformers.utils.fx, which transformers >= 5 no longer ships; describe other "
                "architectures with UniformPartitioner.register_family(matches, blocks, stage_cls) — see its docstring."
                + why)
        assert len(blocks) >= n, "more pipeline stages than transformer blocks"
        costs = [sum(p.numel() for p in blk.parameters()) for blk in blocks]  # embeddings excluded, as in the reference
        b = _balanced_cuts(costs, n)
        if rotary:
            stage_cls = RotaryDecoderStage
        else:
            stage_cls = GPT2Stage if hasattr(model.transformer, "wte") else BloomStage
        return [stage_cls(model, b[i], b[i + 1], is_first=(i == 0), is_last=(i == n - 1)) for i in range(n)]


def _get_partitioner(policy: PartitionPolicy):
    """The partitioner class that implements ``policy``."""
    return {PartitionPolicy.UNIFORM: UniformPartitioner}[policy]


def get_model_partition(module: nn.Module, policy: PartitionPolicy, parallel_context: ParallelContext) -> nn.Module:
    """The stage of ``module`` that belongs to this rank."""
    stages = _get_partitioner(policy)(module, parallel_context).split()
    from pipegoose_b200.nn.pipeline_parallel._utils import get_partition_idx

    return stages[get_partition_idx(parallel_context)]
