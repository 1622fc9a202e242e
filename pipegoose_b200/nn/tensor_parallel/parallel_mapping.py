"""Which sub-modules are column / row / lm-head parallel (parity: reference nn/tensor_parallel/parallel_mapping.py:4-52;
its table has Bloom and Albert and a "make this extendable" note — here more 🤗 families are listed and
``TensorParallelMapping.register`` adds new ones at run time).

Patterns are matched against the last two components of a module path (``model.layers.0.self_attn.q_proj`` ->
``self_attn.q_proj``), the longest matching pattern wins.  The class-swap path gathers column-parallel outputs and
scatters row-parallel inputs, so any model whose projections are ``nn.Linear`` works without touching its attention
code; the sequence-parallel fast path (pipegoose_b200.models) does not use this table.
"""
from typing import List

from pipegoose_b200.nn.parallel_mapping import ParallelInfo, ParallelMapping


class Column(ParallelInfo):
    pass


class Row(ParallelInfo):
    pass


class LMHead(ParallelInfo):
    pass


class TensorParallelMapping(ParallelMapping):
    __MAPPING__ = {
        "bloom-560m": [
            Column(("mlp.dense_h_to_4h", "self_attention.query_key_value")),
            Row(("mlp.dense_4h_to_h", "self_attention.dense")),
            LMHead(("lm_head",)),
        ],
        "albert-base-v2": [
            Column(("attention.query", "attention.key", "attention.value", "ffn")),
            Row(("attention.dense", "ffn_output")),
            LMHead(("predictions.decoder",)),
        ],
        # decoder families with separate q/k/v projections (OPT; LLaMA / Mistral / Qwen2 naming)
        "opt": [
            Column(("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "fc1")),
            Row(("self_attn.out_proj", "fc2")),
            LMHead(("lm_head",)),
        ],
        "llama": [
            Column(("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "mlp.gate_proj", "mlp.up_proj")),
            Row(("self_attn.o_proj", "mlp.down_proj")),
            LMHead(("lm_head",)),
        ],
        "gpt_neox": [
            Column(("attention.query_key_value", "mlp.dense_h_to_4h")),
            Row(("attention.dense", "mlp.dense_4h_to_h")),
            LMHead(("embed_out",)),
        ],
        # encoder: BERT / RoBERTa naming
        "bert": [
            Column(("self.query", "self.key", "self.value", "intermediate.dense")),
            Row(("output.dense",)),
        ],
    }

    @classmethod
    def register(cls, model_key: str, infos: List[ParallelInfo]):
        """Add (or replace) the entry of a model family: ``register("my-model", [Column(("attn.wq", ...)),
        Row(("attn.wo",)), LMHead(("head",))])`` before calling ``TensorParallel(model, ctx).parallelize()``."""
        assert all(isinstance(i, (Column, Row, LMHead)) for i in infos), "entries must be Column / Row / LMHead"
        cls.__MAPPING__[model_key] = list(infos)

    @classmethod
    def unregister(cls, model_key: str):
        cls.__MAPPING__.pop(model_key, None)

    @staticmethod
    def is_column_parallel(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), Column)

    @staticmethod
    def is_row_parallel(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), Row)

    @staticmethod
    def is_lm_head(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), LMHead)
