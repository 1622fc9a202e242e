"""Offline checkpoint conversion (nn/checkpoint_convert.py): shards written by a TP / TP x PP / expert-parallel job are
merged into the state dict of the unsharded model, and re-cut for another tensor-parallel size."""
import copy
import os

import pytest
import torch

from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
from pipegoose_b200.nn import PipelineParallel, TensorParallel
from pipegoose_b200.nn.checkpoint_convert import consolidate_checkpoint, main, reshard_checkpoint
from pipegoose_b200.nn.utils import from_pretrained, save_pretrained
from pipegoose_b200.testing.utils import init_parallel_context, spawn

VOCAB = 90      # not a multiple of 8 * tp: the vocabulary is zero-padded before it is cut


def _model(n_layer=4):
    torch.manual_seed(0)
    return BloomForCausalLM(BloomConfig(vocab_size=VOCAB, hidden_size=32, n_layer=n_layer, n_head=4))


def run_save(rank, world_size, port, tp, pp, ckp_path, state, hf):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, 1)
    if hf:
        from transformers import BloomConfig as HFConfig
        from transformers import BloomForCausalLM as HFBloom

        model = HFBloom(HFConfig(vocab_size=VOCAB, hidden_size=32, n_layer=4, n_head=4))
        model.load_state_dict(state)
        model = TensorParallel(model, ctx, sequence_parallel=False).parallelize()      # class-swap path
    else:
        model = _model()
        model.load_state_dict(state)
        model = TensorParallel(model, ctx).parallelize()
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    ctx.destroy()


@pytest.mark.parametrize("tp,pp,hf", [(2, 1, False), (2, 2, False), (1, 2, False), (2, 1, True), (2, 2, True)])
def test_consolidated_checkpoint_is_the_unsharded_model(tmp_path, tp, pp, hf):
    if hf:
        from transformers import BloomConfig as HFConfig
        from transformers import BloomForCausalLM as HFBloom

        torch.manual_seed(0)
        full = HFBloom(HFConfig(vocab_size=VOCAB, hidden_size=32, n_layer=4, n_head=4))
    else:
        full = _model()
    state = copy.deepcopy(full.state_dict())
    ckpt = str(tmp_path / "ckpt")
    spawn(run_save, world_size=tp * pp, tp=tp, pp=pp, ckp_path=ckpt, state=state, hf=hf)
    assert os.path.exists(os.path.join(ckpt, "pytorch_model_tp_0_pp_0.bin.layout.json"))
    merged = consolidate_checkpoint(ckpt, tp, pp)
    assert set(merged) == set(state), (set(merged) ^ set(state))
    for k, v in state.items():
        assert merged[k].shape == v.shape and torch.equal(merged[k], v), k
    # the merged dict loads into a fresh unsharded model (strict)
    fresh = copy.deepcopy(full)
    with torch.no_grad():
        for p in fresh.parameters():
            p.zero_()
    fresh.load_state_dict(merged)
    # command line: python -m pipegoose_b200.nn.checkpoint_convert ckpt out.bin --tp .. --pp ..
    out = str(tmp_path / "out" / "full.bin")
    main([ckpt, out, "--tp", str(tp), "--pp", str(pp)])
    again = torch.load(out)
    assert all(torch.equal(again[k], v) for k, v in state.items())


def run_load_resharded(rank, world_size, port, tp, ckp_path, state):
    ctx = init_parallel_context(rank, world_size, port, tp, 1, 1)
    want = TensorParallel(_model(), ctx)
    want.module.load_state_dict(state)
    want = want.parallelize()
    got = _model()
    with torch.no_grad():
        for p in got.parameters():
            p.normal_()
    got = TensorParallel(got, ctx).parallelize()
    from_pretrained(got, ckp_path=ckp_path, parallel_context=ctx)
    for (k, a), (_, b) in zip(got.state_dict().items(), want.state_dict().items()):
        assert torch.equal(a, b), k
    ctx.destroy()


def test_reshard_tp2_pp2_checkpoint_for_tp4(tmp_path):
    state = copy.deepcopy(_model().state_dict())
    src, dst = str(tmp_path / "src"), str(tmp_path / "dst")
    spawn(run_save, world_size=4, tp=2, pp=2, ckp_path=src, state=state, hf=False)
    reshard_checkpoint(src, dst, 2, 2, 4)
    spawn(run_load_resharded, world_size=4, tp=4, ckp_path=dst, state=state)
    # and back to one file set of a single rank
    reshard_checkpoint(dst, str(tmp_path / "one"), 4, 1, 1)
    one = torch.load(str(tmp_path / "one" / "pytorch_model_tp_0_pp_0.bin"))
    assert all(torch.equal(one[k], v) for k, v in state.items())


def test_stale_or_incomplete_shards_are_refused(tmp_path):
    state = copy.deepcopy(_model().state_dict())
    ckpt = str(tmp_path / "ckpt")
    spawn(run_save, world_size=2, tp=2, pp=1, ckp_path=ckpt, state=state, hf=False)
    with pytest.raises(FileNotFoundError):
        consolidate_checkpoint(ckpt, 4, 1)
    with pytest.raises(ValueError, match="written for tp=2"):
        consolidate_checkpoint(ckpt, 1, 1)
    shard = torch.load(os.path.join(ckpt, "pytorch_model_tp_1_pp_0.bin"))
    shard["transformer.ln_f.weight"] += 1.0                     # a replicated parameter that differs between ranks
    torch.save(shard, os.path.join(ckpt, "pytorch_model_tp_1_pp_0.bin"))
    with pytest.raises(ValueError, match="differs between tensor-parallel ranks"):
        consolidate_checkpoint(ckpt, 2, 1)
    # without the layout files the name table of TensorParallelMapping decides (vocab_size cuts the padding off)
    for f in os.listdir(ckpt):
        if f.endswith(".layout.json"):
            os.remove(os.path.join(ckpt, f))
    merged = consolidate_checkpoint(ckpt, 2, 1, vocab_size=VOCAB, check_replicas=False)
    for k, v in state.items():
        assert merged[k].shape == v.shape, k
        if k != "transformer.ln_f.weight":
            assert torch.equal(merged[k], v), k


def _moe(ctx, world):
    """Bloom with 4 distinct experts in layers 0 and 2: global expert e of layer li is seeded by (li, e)."""
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.nn.expert_parallel import ExpertParallel, Top2Router

    torch.manual_seed(0)
    model = BloomForCausalLM(BloomConfig(vocab_size=VOCAB, hidden_size=32, n_layer=3, n_head=4))
    model = ExpertParallel(model, 4, mapping=[0, 2], router=Top2Router(None, 4, 32), parallel_context=ctx, fused=False).parallelize()
    tp_rank = ctx.get_local_rank(ParallelMode.TENSOR)
    for li in (0, 2):
        layer = model.transformer.h[li].mlp
        for i, e in enumerate(layer.experts):
            g = torch.Generator().manual_seed(50 + 10 * li + tp_rank * len(layer.experts) + i)
            for p in e.parameters():
                p.data = torch.randn(p.shape, generator=g) * 0.3
    return TensorParallel(model, ctx).parallelize()


def run_moe_save(rank, world_size, port, ckp_path, pp=1):
    ctx = init_parallel_context(rank, world_size, port, world_size // pp, pp, 1)
    model = _moe(ctx, world_size)
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    ctx.destroy()


def test_expert_parallel_checkpoint_renumbers_the_experts(tmp_path):
    """Experts spread over the tensor group are saved under LOCAL indices; the merged dict has them under their global
    indices — the state dict of the same model built on one rank."""
    two, one = str(tmp_path / "two"), str(tmp_path / "one")
    spawn(run_moe_save, world_size=2, ckp_path=two)
    spawn(run_moe_save, world_size=1, ckp_path=one)
    want = torch.load(os.path.join(one, "pytorch_model_tp_0_pp_0.bin"))
    staged = str(tmp_path / "staged")
    spawn(run_moe_save, world_size=4, ckp_path=staged, pp=2)              # experts over TP2, layers over PP2
    for merged in (consolidate_checkpoint(two, 2, 1), consolidate_checkpoint(staged, 2, 2)):
        assert set(merged) == set(want), set(merged) ^ set(want)
        assert any(".experts.3." in k for k in merged)
        for k, v in want.items():
            assert torch.equal(merged[k], v), k


def test_layout_of_a_fused_moe_layer_records_the_expert_dimension():
    """The fused NVLink MoE layer keeps its experts stacked (``[E_local, ...]``): the layout file says "cut along
    dimension 0, E experts in total"; another pipeline stage's zero-size stand-in stays "absent"."""
    from torch import nn

    import pipegoose_b200.ops.moe as moe
    from pipegoose_b200.distributed.parallel_mode import ParallelMode
    from pipegoose_b200.nn.checkpoint_convert import shard_layout, shard_state_dict

    class Ctx:
        tensor_parallel_size, pipeline_parallel_size = 2, 2

        def get_local_rank(self, mode):
            return 1 if mode is ParallelMode.TENSOR else 0

    layer = moe.FusedExpertLayer.__new__(moe.FusedExpertLayer)     # (the constructor needs a process group: fields by hand)
    nn.Module.__init__(layer)
    layer.num_experts = 8
    layer.w1, layer.b1 = nn.Parameter(torch.zeros(4, 16, 4)), nn.Parameter(torch.zeros(4, 16))
    layer.w2, layer.b2 = nn.Parameter(torch.zeros(0)), nn.Parameter(torch.zeros(4, 4))
    root = nn.Module()
    root.mlp = layer
    keys = shard_layout(root, Ctx())["keys"]
    assert keys["mlp.w1"] == {"dim": 0, "full": 8} and keys["mlp.b1"] == {"dim": 0, "full": 8}
    assert keys["mlp.w2"] == {"absent": True}
    # re-cutting such a tensor splits the experts, without the vocabulary padding rule
    full = {"mlp.b1": torch.arange(8 * 16.0).view(8, 16)}
    part = shard_state_dict(full, keys, 4, 3)
    assert torch.equal(part["mlp.b1"], full["mlp.b1"][6:8])


def _tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    return LlamaForCausalLM(LlamaConfig(vocab_size=VOCAB, hidden_size=32, intermediate_size=64, num_hidden_layers=4,
                                        num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=64,
                                        tie_word_embeddings=False))


def run_llama_save(rank, world_size, port, tp, pp, ckp_path, state):
    ctx = init_parallel_context(rank, world_size, port, tp, pp, 1)
    model = _tiny_llama()
    model.load_state_dict(state)
    model = TensorParallel(model, ctx).parallelize()              # class-swap path (q/k/v/o, gate/up/down, untied lm_head)
    if pp > 1:
        model = PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()
    save_pretrained(model, ckp_path=ckp_path, parallel_context=ctx)
    ctx.destroy()


def test_consolidation_is_driven_by_metadata_not_by_bloom_names(tmp_path):
    """A LLaMA-style 🤗 model (other module names, untied lm_head, rotary stages): the layout files carry everything."""
    state = copy.deepcopy(_tiny_llama().state_dict())
    ckpt = str(tmp_path / "ckpt")
    spawn(run_llama_save, world_size=4, tp=2, pp=2, ckp_path=ckpt, state=state)
    merged = consolidate_checkpoint(ckpt, 2, 2)
    assert set(merged) == set(state)
    for k, v in state.items():
        assert merged[k].shape == v.shape and torch.equal(merged[k], v), k


def run_load_resharded_hf(rank, world_size, port, ckp_path, state):
    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    ctx = init_parallel_context(rank, world_size, port, world_size, 1, 1)

    def build(load):
        torch.manual_seed(1)
        model = HFBloom(HFConfig(vocab_size=VOCAB, hidden_size=32, n_layer=4, n_head=4))
        if load:
            model.load_state_dict(state)
        return TensorParallel(model, ctx, sequence_parallel=False).parallelize()

    want, got = build(True), build(False)
    from_pretrained(got, ckp_path=ckp_path, parallel_context=ctx)
    for (k, a), (_, b) in zip(got.state_dict().items(), want.state_dict().items()):
        assert a.shape == b.shape and torch.equal(a, b), k
    ctx.destroy()


def test_reshard_keeps_the_vocabulary_padding_rule_of_the_path_that_wrote_the_checkpoint(tmp_path):
    """The class-swap path pads the vocabulary to a multiple of the group (90 -> 92 at tp 4), the fast path to 8 x group
    (90 -> 96): re-cut shards must load into a job of the SAME path, so the layout records the rule."""
    import json

    from transformers import BloomConfig as HFConfig
    from transformers import BloomForCausalLM as HFBloom

    torch.manual_seed(0)
    state = copy.deepcopy(HFBloom(HFConfig(vocab_size=VOCAB, hidden_size=32, n_layer=4, n_head=4)).state_dict())
    src, dst = str(tmp_path / "src"), str(tmp_path / "dst")
    spawn(run_save, world_size=2, tp=2, pp=1, ckp_path=src, state=state, hf=True)
    lay = json.load(open(os.path.join(src, "pytorch_model_tp_0_pp_0.bin.layout.json")))["keys"]
    assert lay["transformer.word_embeddings.weight"] == {"dim": 0, "full": VOCAB, "vocab": True, "vocab_multiple": 1}
    reshard_checkpoint(src, dst, 2, 1, 4)
    assert torch.load(os.path.join(dst, "pytorch_model_tp_3_pp_0.bin"))["transformer.word_embeddings.weight"].shape[0] == 23
    spawn(run_load_resharded_hf, world_size=4, ckp_path=dst, state=state)


def run_load_into_pipeline(rank, world_size, port, ckp_path, state):
    ctx = init_parallel_context(rank, world_size, port, 2, 2, 1)

    def build(load):
        model = _model()
        if load:
            model.load_state_dict(state)
        else:
            with torch.no_grad():
                for p in model.parameters():
                    p.normal_()
        model = TensorParallel(model, ctx).parallelize()
        return PipelineParallel(model, num_microbatches=2, parallel_context=ctx).parallelize()

    want, got = build(True), build(False)
    from_pretrained(got, ckp_path=ckp_path, parallel_context=ctx)
    n = 0
    for (k, a), (_, b) in zip(got.state_dict().items(), want.state_dict().items()):
        assert a.shape == b.shape and torch.equal(a, b), k
        n += a.numel() > 0
    assert n > 0
    ctx.destroy()


def test_reshard_for_a_pipelined_job(tmp_path):
    """TP2 checkpoint -> TP2 x PP2 job: every pipeline rank's file holds the whole tensor-sliced model, a stage loads what it
    owns (its ``_pg_pipeline_stage.`` aliases share that storage) and ignores the other stages' tensors."""
    state = copy.deepcopy(_model().state_dict())
    src, dst = str(tmp_path / "src"), str(tmp_path / "dst")
    spawn(run_save, world_size=2, tp=2, pp=1, ckp_path=src, state=state, hf=False)
    main([src, dst, "--tp", "2", "--pp", "1", "--new-tp", "2", "--new-pp", "2"])
    assert sorted(f for f in os.listdir(dst) if f.endswith(".bin")) == [
        "pytorch_model_tp_0_pp_0.bin", "pytorch_model_tp_0_pp_1.bin", "pytorch_model_tp_1_pp_0.bin", "pytorch_model_tp_1_pp_1.bin"]
    spawn(run_load_into_pipeline, world_size=4, ckp_path=dst, state=state)
    merged = consolidate_checkpoint(dst, 2, 2)                    # and such a directory merges back like any other
    assert all(torch.equal(merged[k], v) for k, v in state.items())
