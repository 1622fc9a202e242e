"""Name-based lookup of which sub-modules a wrapper touches (parity: reference nn/parallel_mapping.py:4-37)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple


class ParallelInfo:
    def __init__(self, module_name: Tuple[str, ...], **kwargs):
        self.module_name = module_name
        self.kwargs = kwargs


class ParallelMapping:
    """Sub-classes define ``__MAPPING__ = {model_key: [ParallelInfo, ...]}``.

    ``_search(name)`` reduces a dotted module path to its last two components
    (``transformer.h.0.mlp.dense_h_to_4h`` -> ``mlp.dense_h_to_4h``) and returns the first
    ``ParallelInfo`` that lists a pattern contained in it, over every model key.
    """

    __MAPPING__: Dict[str, List[ParallelInfo]] = {}

    @staticmethod
    def _extract_module_name(module_name: str) -> str:
        parts = module_name.split(".")
        return ".".join(parts[-2:]) if len(parts) >= 2 else module_name

    @classmethod
    def _search(cls, module_name: str) -> Optional[ParallelInfo]:
        tail = cls._extract_module_name(module_name)
        best, best_len = None, -1
        for infos in cls.__MAPPING__.values():
            for info in infos:
                for pattern in info.module_name:
                    # longest matching pattern wins ("ffn_output" beats "ffn")
                    if pattern in tail and len(pattern) > best_len:
                        best, best_len = info, len(pattern)
        return best
