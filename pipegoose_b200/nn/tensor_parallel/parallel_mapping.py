"""Which Bloom (and Albert) sub-modules are column / row / lm-head parallel
(parity: reference nn/tensor_parallel/parallel_mapping.py:4-52)."""
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
    }

    @staticmethod
    def is_column_parallel(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), Column)

    @staticmethod
    def is_row_parallel(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), Row)

    @staticmethod
    def is_lm_head(module_name: str) -> bool:
        return isinstance(TensorParallelMapping._search(module_name), LMHead)
