"""Offline checkpoint conversion: merge the per-(tp, pp) shard files that ``nn.utils.save_pretrained`` writes into ONE
unsharded state dict (what 🤗 ``load_state_dict`` / ``models.bloom.BloomForCausalLM`` read), and cut such a state
dict for another tensor-parallel size — train at TP2 x PP2, serve at TP1 or TP4 — without starting a distributed job.

The reference stops at writing the shards (nn/utils.py:33-50; ``deparallelize`` is unimplemented,
nn/tensor_parallel/tensor_parallel.py:79-82).  ``TensorParallel.deparallelize()`` here is the ONLINE inverse (a
collective on a live job); this module is the OFFLINE one (plain files, one process, CPU).

How a key was cut is not guessed from shapes: ``save_pretrained`` writes a layout file next to every shard
(``<shard>.layout.json``, from the ``ParallelMetadata`` the parallelizers attach to each sliced parameter):

* ``{"dim": d, "full": n}`` — split along ``d`` over the tensor group; ``n`` is the unsharded size (``"vocab": true``:
  a vocabulary table, zero-padded to a multiple of the group before slicing — the padding is cut off again); the
  stacked ``[E_local, ...]`` tensors of a fused MoE layer are recorded the same way (``d = 0``, ``n`` experts);
* ``{"expert": g}`` — the key belongs to local expert ``l`` of a layer whose experts are spread over the tensor group;
  ``g`` is its global index and the consolidated key is renumbered to it;
* ``{"absent": true}`` — a parameter of another pipeline stage (a stage keeps zero-size stand-ins);
* no entry — replicated over the tensor group (LayerNorms, row-parallel biases, routers): rank 0's copy is taken and
  the other ranks' copies are compared with it (``check_replicas``).

Checkpoints written before the layout files existed fall back to the name rules of ``TensorParallelMapping``.
"""
from __future__ import annotations

import json
import os
import re
from pathlib import Path
from typing import Dict, Optional

import torch
from torch import nn

from pipegoose_b200.constants import CHECKPOINT_WEIGHTS_NAME

LAYOUT_SUFFIX = ".layout.json"
_STAGE_ALIAS = "_pg_pipeline_stage."      # a pipeline wrapper's second handle on its stage's parameters
_EXPERT_KEY = re.compile(r"^(.*\.experts\.)(\d+)(\..*)$")


# ------------------------------------------------------------------------------------------------------------------
# layout of a live (parallelized) module — written by save_pretrained
# ------------------------------------------------------------------------------------------------------------------
def shard_layout(module: nn.Module, parallel_context) -> Dict:
    """``{"tp": .., "pp": .., "tp_rank": .., "pp_rank": .., "keys": {state-dict key: entry}}`` for this rank's shard."""
    from pipegoose_b200.distributed.parallel_mode import ParallelMode

    keys: Dict[str, Dict] = {}
    tp_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    # experts spread over the tensor group: local index -> global index
    expert_prefix: Dict[str, int] = {}
    try:
        from pipegoose_b200.nn.expert_parallel.experts import Experts
    except Exception:  # pragma: no cover
        Experts = ()
    for name, mod in module.named_modules():
        if isinstance(mod, Experts) and getattr(mod, "sharded", False) and parallel_context.tensor_parallel_size > 1:
            first = tp_rank * mod.num_local_experts
            for local in range(len(mod.experts)):
                expert_prefix[f"{name}.experts.{local}." if name else f"experts.{local}."] = first + local
    stacked: Dict[str, int] = {}
    try:
        from pipegoose_b200.ops.moe import FusedExpertLayer
    except Exception:  # pragma: no cover
        FusedExpertLayer = ()
    for name, mod in module.named_modules():
        if isinstance(mod, FusedExpertLayer) and parallel_context.tensor_parallel_size > 1:
            for leaf in ("w1", "b1", "w2", "b2"):       # [E_local, ...]: the expert dimension is the one that is cut
                stacked[f"{name}.{leaf}" if name else leaf] = int(mod.num_experts)
    state = dict(module.named_parameters(remove_duplicate=False))
    state.update(dict(module.named_buffers(remove_duplicate=False)))
    for key, t in state.items():
        entry: Dict = {}
        meta = getattr(t, "parallel_metadata", None)
        if t.numel() == 0 and t.dim() == 1:
            entry["absent"] = True
        elif meta is not None and meta.is_sliced and meta.partition_dim is not None:
            entry["dim"] = int(meta.partition_dim)
            if meta.full_size is not None:
                entry["full"] = int(meta.full_size)
            if getattr(meta, "is_vocab", False):
                entry["vocab"] = True
                entry["vocab_multiple"] = int(getattr(meta, "vocab_multiple", 1))
        if key in stacked and not entry.get("absent"):
            entry = {"dim": 0, "full": stacked[key]}
        for prefix, g in expert_prefix.items():
            if key.startswith(prefix):
                entry["expert"] = g
                break
        if entry:
            keys[key] = entry
    return {"tp": parallel_context.tensor_parallel_size, "pp": parallel_context.pipeline_parallel_size,
            "tp_rank": tp_rank, "pp_rank": parallel_context.get_local_rank(ParallelMode.PIPELINE), "keys": keys}


def write_layout(module: nn.Module, shard_path: str, parallel_context) -> None:
    tmp = f"{shard_path}{LAYOUT_SUFFIX}.tmp.{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump(shard_layout(module, parallel_context), f)
    os.replace(tmp, shard_path + LAYOUT_SUFFIX)


# ------------------------------------------------------------------------------------------------------------------
# consolidate
# ------------------------------------------------------------------------------------------------------------------
def _name_rule(key: str) -> Optional[Dict]:
    """Fallback for checkpoints without layout files: the tensor-parallel name table."""
    from pipegoose_b200.nn.tensor_parallel.parallel_mapping import TensorParallelMapping as M
    from pipegoose_b200.nn.tensor_parallel.parallelizer import EmbeddingParallelizer

    module_name, _, leaf = key.rpartition(".")
    if M.is_column_parallel(module_name):
        return {"dim": 0}
    if M.is_row_parallel(module_name):
        return {"dim": 1} if leaf == "weight" else None
    last = module_name.rsplit(".", 1)[-1]
    if M.is_lm_head(module_name) or "word_embeddings" == last or last in EmbeddingParallelizer.TOKEN_EMBEDDING_NAMES:
        return {"dim": 0}
    return None


def _read_shards(ckp_path: str, tp: int, pp: int, ckp_name: str):
    shards, layouts = {}, {}
    for p in range(pp):
        for t in range(tp):
            path = os.path.join(ckp_path, ckp_name.format(t, p))
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} is missing: tensor_parallel_size={tp} x pipeline_parallel_size={pp} needs "
                                        f"{tp * pp} shard files")
    for p in range(pp):
        for t in range(tp):
            path = os.path.join(ckp_path, ckp_name.format(t, p))
            shards[t, p] = torch.load(path, map_location="cpu")
            lay = path + LAYOUT_SUFFIX
            if os.path.exists(lay):
                with open(lay) as f:
                    layouts[t, p] = json.load(f)
                got = (layouts[t, p]["tp"], layouts[t, p]["pp"])
                if got != (tp, pp):
                    raise ValueError(f"{lay} was written for tp={got[0]} pp={got[1]}, asked to consolidate tp={tp} pp={pp}")
    return shards, layouts


def consolidate_checkpoint(ckp_path: str, tensor_parallel_size: int, pipeline_parallel_size: int = 1,
                           ckp_name: str = CHECKPOINT_WEIGHTS_NAME, vocab_size: Optional[int] = None,
                           check_replicas: bool = True) -> Dict[str, torch.Tensor]:
    """The unsharded state dict of the checkpoint under ``ckp_path``.  ``vocab_size``: only needed for checkpoints
    without layout files whose vocabulary was zero-padded (the layout files carry the true size)."""
    tp, pp = tensor_parallel_size, pipeline_parallel_size
    shards, layouts = _read_shards(ckp_path, tp, pp, ckp_name)
    have_layout = len(layouts) == tp * pp
    out: Dict[str, torch.Tensor] = {}
    for p in range(pp):
        ref = shards[0, p]
        for key in ref:
            if key.startswith(_STAGE_ALIAS) or ("." + _STAGE_ALIAS) in key:
                continue
            entries = [layouts[t, p]["keys"].get(key, {}) if have_layout else None for t in range(tp)]
            parts = [shards[t, p][key] for t in range(tp)]
            if have_layout:
                if entries[0].get("absent"):
                    continue
                rule = entries[0]
            else:
                if parts[0].numel() == 0 and parts[0].dim() == 1:
                    continue        # another stage's parameter
                rule = _name_rule(key) or {}
                if vocab_size is not None and rule.get("dim") == 0 and parts[0].dim() == 2 and \
                        parts[0].shape[0] * tp >= vocab_size > parts[0].shape[0] * (tp - 1) and _is_vocab_key(key):
                    rule = dict(rule, full=vocab_size)
            if "expert" in rule:
                # experts spread over the group: every rank contributes ITS experts under their global numbers
                for t in range(tp):
                    g = entries[t]["expert"]
                    m = _EXPERT_KEY.match(key)
                    assert m is not None, key
                    _put(out, f"{m.group(1)}{g}{m.group(3)}", parts[t].clone())
                continue
            if "dim" in rule and tp > 1:
                full = torch.cat(parts, dim=rule["dim"])
                if rule.get("full") is not None and full.shape[rule["dim"]] != rule["full"]:
                    full = full.narrow(rule["dim"], 0, rule["full"]).contiguous()
                _put(out, key, full)
                continue
            if check_replicas:
                for t in range(1, tp):
                    if parts[t].shape != parts[0].shape or not torch.equal(parts[t], parts[0]):
                        raise ValueError(f"{key} differs between tensor-parallel ranks 0 and {t} of pipeline stage {p} but "
                                         "is not recorded as sliced: stale shard files from different runs?")
            _put(out, key, parts[0].clone())
    return out


def _is_vocab_key(key: str) -> bool:
    from pipegoose_b200.nn.tensor_parallel.parallel_mapping import TensorParallelMapping as M
    from pipegoose_b200.nn.tensor_parallel.parallelizer import EmbeddingParallelizer

    module_name = key.rpartition(".")[0]
    last = module_name.rsplit(".", 1)[-1]
    return M.is_lm_head(module_name) or last == "word_embeddings" or last in EmbeddingParallelizer.TOKEN_EMBEDDING_NAMES


def _put(out: Dict[str, torch.Tensor], key: str, value: torch.Tensor) -> None:
    """A key two pipeline stages hold (the tied embedding / lm_head table on the first and last stage) must agree."""
    old = out.get(key)
    if old is not None and (old.shape != value.shape or not torch.equal(old, value)):
        raise ValueError(f"{key} is held by two pipeline stages with different values (tied parameters are kept in step by "
                         "the pipeline engine: was the checkpoint written in the middle of a step?)")
    out[key] = value


# ------------------------------------------------------------------------------------------------------------------
# reshard
# ------------------------------------------------------------------------------------------------------------------
def shard_state_dict(full: Dict[str, torch.Tensor], layout_keys: Dict[str, Dict], tensor_parallel_size: int, tp_rank: int,
                     vocab_multiple: int = 8) -> Dict[str, torch.Tensor]:
    """Rank ``tp_rank``'s shard of an unsharded state dict for a tensor group of ``tensor_parallel_size``:
    ``layout_keys[key] = {"dim": d[, "full": n]}`` says which keys are cut (the ``keys`` of any rank's layout file of the
    source checkpoint works: the SET of sliced keys does not depend on the group size).  Keys recorded with ``vocab``
    (vocabulary tables, lm_head rows / bias) are zero-padded to a multiple of ``vocab_multiple * tensor_parallel_size`` first:
    the multiple recorded in the layout (8 for the sequence-parallel fast path, 1 for the class-swap path), else the argument."""
    tp = tensor_parallel_size
    out = {}
    for key, t in full.items():
        rule = layout_keys.get(key, {})
        if "expert" in rule:
            raise ValueError("expert-sharded checkpoints are re-cut by ExpertParallel on a consolidated model, not offline")
        if "dim" not in rule or tp == 1:
            out[key] = t.clone()
            continue
        d = rule["dim"]
        if rule.get("vocab") and d == 0:
            mult = int(rule.get("vocab_multiple", vocab_multiple)) * tp     # the rule of the path that wrote the checkpoint
            padded = (t.shape[0] + mult - 1) // mult * mult
            if padded != t.shape[0]:
                t = torch.cat([t, t.new_zeros(padded - t.shape[0], *t.shape[1:])], dim=0)
        if t.shape[d] % tp:
            raise ValueError(f"{key}: dimension {d} of size {t.shape[d]} does not divide by {tp}")
        w = t.shape[d] // tp
        out[key] = t.narrow(d, tp_rank * w, w).clone().contiguous()
    return out


def reshard_checkpoint(src: str, dst: str, tensor_parallel_size: int, pipeline_parallel_size: int,
                       new_tensor_parallel_size: int, ckp_name: str = CHECKPOINT_WEIGHTS_NAME,
                       vocab_multiple: int = 8, new_pipeline_parallel_size: int = 1) -> None:
    """Rewrite the checkpoint under ``src`` (tp x pp shards) as ``new_tensor_parallel_size`` x
    ``new_pipeline_parallel_size`` shards under ``dst`` for ``from_pretrained`` of a job with that layout.  Every pipeline
    rank's file holds the whole (tensor-sliced) model — which blocks a stage owns is the partitioner's decision at run
    time; ``from_pretrained`` takes what its stage holds and ignores the rest."""
    _, layouts = _read_shards(src, tensor_parallel_size, pipeline_parallel_size, ckp_name)
    if len(layouts) != tensor_parallel_size * pipeline_parallel_size:
        raise ValueError("resharding needs the layout files save_pretrained writes next to the shards")
    full = consolidate_checkpoint(src, tensor_parallel_size, pipeline_parallel_size, ckp_name)
    keys: Dict[str, Dict] = {}
    for lay in layouts.values():
        for k, e in lay["keys"].items():
            if not e.get("absent") and not k.startswith(_STAGE_ALIAS):
                keys.setdefault(k, e)
    Path(dst).mkdir(parents=True, exist_ok=True)
    for r in range(new_tensor_parallel_size):
        shard = shard_state_dict(full, keys, new_tensor_parallel_size, r, vocab_multiple)
        for p in range(new_pipeline_parallel_size):
            path = os.path.join(dst, ckp_name.format(r, p))
            torch.save(shard, path)
            with open(path + LAYOUT_SUFFIX, "w") as f:
                json.dump({"tp": new_tensor_parallel_size, "pp": new_pipeline_parallel_size, "tp_rank": r, "pp_rank": p,
                           "keys": {k: e for k, e in keys.items() if "dim" in e}}, f)


def main(argv=None) -> None:
    import argparse

    ap = argparse.ArgumentParser(description="merge / re-cut pipegoose_b200 checkpoint shards offline")
    ap.add_argument("src")
    ap.add_argument("dst", help="a .bin/.pt file (consolidated state dict) or, with --new-tp, a directory of shards")
    ap.add_argument("--tp", type=int, required=True)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--new-tp", type=int, default=0)
    ap.add_argument("--new-pp", type=int, default=1)
    ap.add_argument("--vocab-size", type=int, default=None)
    a = ap.parse_args(argv)
    if a.new_tp:
        reshard_checkpoint(a.src, a.dst, a.tp, a.pp, a.new_tp, new_pipeline_parallel_size=a.new_pp)
        print(f"wrote {a.new_tp} x {a.new_pp} shard(s) to {a.dst}")
    else:
        sd = consolidate_checkpoint(a.src, a.tp, a.pp, vocab_size=a.vocab_size)
        Path(os.path.dirname(os.path.abspath(a.dst))).mkdir(parents=True, exist_ok=True)
        torch.save(sd, a.dst)
        print(f"wrote {len(sd)} tensors ({sum(t.numel() for t in sd.values()):,} elements) to {a.dst}")


if __name__ == "__main__":  # python -m pipegoose_b200.nn.checkpoint_convert ckpt/ full.bin --tp 2 --pp 2
    main()
