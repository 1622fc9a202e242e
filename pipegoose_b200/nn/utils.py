"""Checkpoint I/O (parity: reference nn/utils.py:11-50): one file per (tp_rank, pp_rank) named
``pytorch_model_tp_{tp}_pp_{pp}.bin``.  Only data-parallel rank 0 writes (the reference has every
replica write the same path) and the directory is created when missing.  Next to each shard goes
``<shard>.layout.json`` (which keys were cut along which dimension), which ``nn/checkpoint_convert.py`` uses to
merge the shards into one state dict or to re-cut them for another tensor-parallel size, offline."""
from __future__ import annotations

import os
from pathlib import Path

import torch
from torch import nn

from pipegoose_b200.constants import CHECKPOINT_PATH_NAME, CHECKPOINT_WEIGHTS_NAME
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def _ckpt_file(ckp_path: str, ckp_name: str, parallel_context: ParallelContext) -> str:
    tp_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    pp_rank = parallel_context.get_local_rank(ParallelMode.PIPELINE)
    return os.path.join(ckp_path, ckp_name.format(tp_rank, pp_rank))


def _atomic_save(obj, path: str):
    """Write next to the target and rename: a job killed mid-write never leaves a truncated checkpoint behind (the
    previous complete file stays in place until the new one is whole)."""
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def from_pretrained(module: nn.Module, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                    ckp_name: str = CHECKPOINT_WEIGHTS_NAME, strict: bool = True):
    """Load this rank's (tp, pp) shard into an already parallelized ``module``.  ``strict`` (default, as the
    reference's ``load_state_dict``): every parameter and buffer of the module must be in the checkpoint."""
    path = _ckpt_file(ckp_path, ckp_name, parallel_context)
    if not os.path.exists(path):
        raise ValueError(f"ckp_path {path} does not exist")
    state_dict = torch.load(path, map_location="cpu")
    with torch.no_grad():
        own = module.state_dict()
        # strict, like the reference's ``module.load_state_dict``: a stale or partial shard (another model config,
        # parallel layout or MoE mapping) must not resume silently with randomly initialised leftovers
        unexpected = [k for k in state_dict if k not in own]
        if unexpected:
            raise KeyError(f"unexpected key(s) {unexpected[:5]} in checkpoint {path}")
        # a pipelined module lists its stage's parameters a second time under ``_pg_pipeline_stage.`` (the same storage):
        # shards re-cut offline (nn/checkpoint_convert.py) only carry the model's own names
        missing = [k for k in own if k not in state_dict and not k.startswith("_pg_pipeline_stage.")]
        if missing and strict:
            raise KeyError(f"checkpoint {path} lacks {len(missing)} parameter(s) / buffer(s) of this module, e.g. "
                           f"{missing[:5]}")

        def foreign(k, v):
            """Another pipeline stage's parameter: this rank keeps a zero-size stand-in, a shard written for every stage
            alike (re-cut offline) holds the tensor — nothing to load here."""
            return own[k].numel() == 0 and own[k].dim() == 1 and v.numel() > 0

        for k, v in state_dict.items():
            if tuple(own[k].shape) != tuple(v.shape) and not foreign(k, v):
                raise ValueError(f"shape mismatch for {k} in {path}: checkpoint {tuple(v.shape)}, module "
                                 f"{tuple(own[k].shape)} (different model config or parallel layout?)")
        for k, v in state_dict.items():
            if not foreign(k, v):
                own[k].copy_(v)
    return module


def save_pretrained(module: nn.Module, ckp_name: str = CHECKPOINT_WEIGHTS_NAME, ckp_path: str = CHECKPOINT_PATH_NAME,
                    parallel_context: ParallelContext = None):
    """Save this rank's (tp, pp) shard; data-parallel replicas > 0 skip the write."""
    Path(ckp_path).mkdir(parents=True, exist_ok=True)
    if parallel_context.get_local_rank(ParallelMode.DATA) == 0:
        state = {k: v.detach().cpu() for k, v in module.state_dict().items()}
        path = _ckpt_file(ckp_path, ckp_name, parallel_context)
        _atomic_save(state, path)
        # how every key was cut, for the offline tools (nn/checkpoint_convert.py: consolidate / reshard)
        from pipegoose_b200.nn.checkpoint_convert import write_layout

        write_layout(module, path, parallel_context)
    if parallel_context.get_world_size(ParallelMode.DATA) > 1:
        import torch.distributed as dist

        dist.barrier(group=parallel_context.get_group(ParallelMode.DATA))


# ------------------------------------------------------------------------------------------------
# resume: optimizer shards + RNG + step (the reference only saves model weights)
# ------------------------------------------------------------------------------------------------
def _optim_file(ckp_path: str, parallel_context: ParallelContext) -> str:
    from pipegoose_b200.constants import CHECKPOINT_OPTIM_NAME

    return os.path.join(ckp_path, CHECKPOINT_OPTIM_NAME.format(
        parallel_context.get_local_rank(ParallelMode.TENSOR), parallel_context.get_local_rank(ParallelMode.PIPELINE),
        parallel_context.get_local_rank(ParallelMode.DATA)))


def _layout(parallel_context: ParallelContext) -> dict:
    return {"tp": parallel_context.tensor_parallel_size, "pp": parallel_context.pipeline_parallel_size,
            "dp": parallel_context.data_parallel_size}


def save_training_state(optim, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                        step: int = 0, extra: dict = None):
    """Every rank writes ITS optimizer shard (ZeRO-1: 1/dp of the fp32 master weights and moments) as
    ``optimizer_tp_{tp}_pp_{pp}_dp_{dp}.bin`` together with the step counter, the RNG states and the parallel
    layout the shard belongs to."""
    Path(ckp_path).mkdir(parents=True, exist_ok=True)
    sd = optim.state_dict()

    def to_cpu(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu()
        if isinstance(x, dict):
            return {k: to_cpu(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(to_cpu(v) for v in x)
        return x

    blob = {"optimizer": to_cpu(sd), "step": int(step), "layout": _layout(parallel_context),
            "rng": capture_rng_state(), "extra": extra or {}, "flat_index": _flat_index(optim)}
    _atomic_save(blob, _optim_file(ckp_path, parallel_context))


def _fused_adam_of(optim):
    """The :class:`FusedAdam` behind ``optim`` (itself, or wrapped by a ``DistributedOptimizer``), else None."""
    from pipegoose_b200.optim.fused_adam import FusedAdam

    inner = getattr(optim, "optim", optim)
    return inner if isinstance(inner, FusedAdam) else None


def _flat_index(optim):
    """``[(offset, numel)]`` of every parameter in the optimizer's flat buffer, in ``param_groups`` order (None for
    optimizers without a flat state).  The flat layout depends on the data-parallel size (region padding), the ORDER of
    the parameters does not: this table is what lets a checkpoint be re-cut for another number of replicas."""
    fa = _fused_adam_of(optim)
    if fa is None:
        return None
    flat = fa.ensure_flat()
    # (None: a parameter the flat state does not hold — another pipeline stage's zero-size stand-in)
    return [tuple(int(x) for x in flat.offsets[id(p)]) if id(p) in flat.offsets else None
            for g in fa.param_groups for p in g["params"]]


def capture_rng_state() -> dict:
    """Every generator a training loop may draw from: torch CPU + CUDA, Python ``random``, numpy."""
    import random

    state = {"torch": torch.get_rng_state(),
             "cuda": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
             "python": random.getstate()}
    try:
        import numpy as np

        state["numpy"] = np.random.get_state()
    except Exception:  # pragma: no cover - numpy missing
        state["numpy"] = None
    return state


def restore_rng_state(state: dict):
    import random

    torch.set_rng_state(state["torch"])
    if state.get("cuda") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state(state["cuda"])
    if state.get("python") is not None:
        random.setstate(state["python"])
    if state.get("numpy") is not None:
        import numpy as np

        np.random.set_state(state["numpy"])


def load_training_state(optim, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                        restore_rng: bool = True) -> dict:
    """Restore what :func:`save_training_state` wrote; returns ``{"step": ..., "extra": ...}`` (plus ``"rng"``, the saved
    generator states, when ``restore_rng=False`` leaves restoring them to the caller)."""
    path = _optim_file(ckp_path, parallel_context)
    if not os.path.exists(path):
        elastic = _load_for_another_dp_size(optim, ckp_path, parallel_context, restore_rng)
        if elastic is not None:
            return elastic
        raise ValueError(f"optimizer checkpoint {path} does not exist")
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob["layout"] != _layout(parallel_context):
        elastic = _load_for_another_dp_size(optim, ckp_path, parallel_context, restore_rng)
        if elastic is not None:
            return elastic
        raise ValueError(f"checkpoint was written for layout {blob['layout']}, this job runs {_layout(parallel_context)}: "
                         "optimizer shards are tied to the parallel layout (only the data-parallel size of a FusedAdam / "
                         "ZeRO-1 checkpoint can change)")
    optim.load_state_dict(blob["optimizer"])
    meta = {"step": blob["step"], "extra": blob["extra"]}
    if restore_rng:
        restore_rng_state(blob["rng"])
    else:
        meta["rng"] = blob["rng"]   # the caller restores it (the Trainer: after it replayed the consumed batches)
    return meta


# ------------------------------------------------------------------------------------------------
# elastic resume: a ZeRO-1 (FusedAdam) checkpoint written by ``dp_old`` replicas, loaded by ``dp_new``
# ------------------------------------------------------------------------------------------------
def _load_for_another_dp_size(optim, ckp_path: str, parallel_context: ParallelContext, restore_rng: bool):
    """The checkpoint has no shard for this rank's ``(tp, pp, dp)`` coordinates, or was written for another layout.  If
    only the DATA-parallel size differs and the optimizer is a (ZeRO-1) ``FusedAdam``, every rank reads all old replicas'
    shards of its ``(tp, pp)`` coordinates and cuts its own slices out of them; returns the metadata, or None when this
    does not apply."""
    import glob
    import re

    from pipegoose_b200.constants import CHECKPOINT_OPTIM_NAME

    if _fused_adam_of(optim) is None:
        return None
    tp_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    pp_rank = parallel_context.get_local_rank(ParallelMode.PIPELINE)
    pattern = os.path.join(ckp_path, CHECKPOINT_OPTIM_NAME.format(tp_rank, pp_rank, "*"))
    found = {}
    for f in glob.glob(pattern):
        m = re.search(r"_dp_(\d+)\.bin$", f)
        if m:
            found[int(m.group(1))] = f
    if not found:
        return None
    first = torch.load(found[min(found)], map_location="cpu", weights_only=False)
    old, now = first["layout"], _layout(parallel_context)
    if (old["tp"], old["pp"]) != (now["tp"], now["pp"]) or old["dp"] == now["dp"]:
        return None
    if sorted(found) != list(range(old["dp"])):
        raise ValueError(f"{ckp_path}: the checkpoint was written by {old['dp']} replicas, found optimizer shards of "
                         f"replicas {sorted(found)} for tp={tp_rank} pp={pp_rank}")
    if not (isinstance(first.get("optimizer"), dict) and "segments" in first["optimizer"]):
        return None      # not a FusedAdam / ZeRO-1 state (e.g. DiLoCo workers: every replica's state is its own)
    if first.get("flat_index") is None:
        raise ValueError("this checkpoint predates the per-parameter index that re-cutting for another data-parallel size "
                         "needs (save it again with this version at the old size)")
    blobs = [first] + [torch.load(found[d], map_location="cpu", weights_only=False) for d in range(1, old["dp"])]
    old_index = [tuple(x) if x is not None else None for x in first["flat_index"]]

    def provider(new_segments, new_index, new_numel):
        return reshard_fused_state([b["optimizer"] for b in blobs], old_index, new_segments, new_index, new_numel)

    load = getattr(optim, "load_resharded_state", None)
    if load is not None:
        load(provider)                       # DistributedOptimizer: now, or when its ZeRO-1 slices are laid out
    else:
        fa = _fused_adam_of(optim)
        fa._lazy_init()
        fa.load_state_dict(provider(list(fa._segments), _flat_index(optim), fa.flat.numel))
    mine = blobs[parallel_context.get_local_rank(ParallelMode.DATA) % old["dp"]]
    meta = {"step": mine["step"], "extra": mine["extra"], "resharded_from_dp": old["dp"]}
    if restore_rng:
        restore_rng_state(mine["rng"])       # (a replica's random stream: the closest there is with another replica count)
    else:
        meta["rng"] = mine["rng"]
    return meta


def reshard_fused_state(old_states: list, old_index: list, new_segments: list, new_index: list, new_numel: int) -> dict:
    """Cut a FusedAdam state dict for ``new_segments`` (ranges of the NEW flat buffer) out of the state dicts of all old
    replicas.  Parameters are matched by their position in ``param_groups`` (``*_index[i] = (offset, numel)`` in the
    old / new flat buffer); every element of every parameter must be owned by exactly one old replica."""
    assert len(old_index) == len(new_index), "the optimizers hold different numbers of parameters"
    pairs = []
    for i, (o, n) in enumerate(zip(old_index, new_index)):
        if (o is None) != (n is None):
            raise ValueError(f"parameter {i} is held by this rank in one of the two jobs only (another pipeline layout?)")
        if o is None:
            continue        # another pipeline stage's parameter in both jobs
        if o[1] != n[1]:
            raise ValueError(f"parameter {i} has {o[1]} elements in the checkpoint and {n[1]} in this model")
        pairs.append((tuple(o), tuple(n)))
    old_numel = max(e for sd in old_states for _, e in sd["segments"])
    out = {"step": old_states[0]["step"], "segments": [tuple(x) for x in new_segments],
           "param_groups": old_states[0].get("param_groups", [])}
    for key in ("master", "exp_avg", "exp_avg_sq"):
        full = torch.full((old_numel,), float("nan"), dtype=torch.float32)
        for sd in old_states:
            off = 0
            for s, e in sd["segments"]:
                full[s:e] = sd[key][off:off + e - s].float()
                off += e - s
        new_full = torch.zeros(new_numel, dtype=torch.float32)
        for (o_old, n), (o_new, _) in pairs:
            piece = full[o_old:o_old + n]
            if torch.isnan(piece).any():
                raise ValueError("the old replicas' slices do not cover every parameter (incomplete checkpoint?)")
            new_full[o_new:o_new + n] = piece
        out[key] = torch.cat([new_full[s:e] for s, e in new_segments]) if new_segments else new_full[:0]
    return out
